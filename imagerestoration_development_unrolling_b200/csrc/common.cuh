// common.cuh - shared helpers for the GLR/GTV kernels (sm_100a).
//
// Every kernel in this directory is written against blockDim-agnostic strided loops
// (`for (i = threadIdx.x; i < n; i += blockDim.x)`) with __syncthreads() only BETWEEN loops and
// block reductions only through block_sum().  That style is what lets tests/emu compile the very
// same sources with g++ (-DGLRGTV_EMU: one "thread" per block, blocks run sequentially) and check
// the index arithmetic and border rules on a machine without a GPU.  The emulation build is test
// infrastructure: the product library is always the nvcc build.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <math.h>

#include "../../include/glrgtv.h"

#ifdef GLRGTV_EMU
// ------------------------------------------------------------------ CPU emulation shim (tests only)
#include <algorithm>
#include <cstring>
#include <cstdlib>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
    emu_dim3() {}
    emu_dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef emu_dim3 dim3;
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local float* emu_smem;  // 256 KB scratch playing the role of dynamic shared memory
// fiber mode (GLR_LAUNCH_FIBERS): every CUDA thread of a block is a cooperative fiber, so __syncthreads() and the
// warp shuffles have their real meaning.  Outside fiber mode a block is one thread and the barrier is a no-op.
extern thread_local bool emu_fiber_mode;
void emu_barrier();
float emu_shfl(float v, int src_lane);   // value of `v` held by lane `src_lane` of the caller's warp
void emu_run_block(unsigned nthreads, void (*fn)(void*), void* arg);
// Asynchronous copies (cp.async / cp.async.bulk) complete at ISSUE time by default - the earliest legal moment, the worst case for
// write-after-read hazards.  With GLRGTV_EMU_ASYNC=late they complete at the LATEST legal moment instead: a cp.async lands when its
// thread executes the wait_group that covers it, a bulk copy when some thread first waits on its mbarrier - the worst case for
// read-before-arrival hazards (a missing or too-shallow wait).  A correct kernel gives the same result in both modes.
extern bool emu_async_late;
void emu_async_push(float* dst, const float* src, int nfloats);
void emu_async_commit();
void emu_async_wait(int keep_groups);           // complete all but the `keep_groups` most recent commit groups of this thread
void emu_bulk_push(float* dst, const float* src, unsigned bytes, const void* bar);
void emu_bulk_wait(const void* bar);
void emu_async_reset_for_launch();
static inline void __syncthreads() { if (emu_fiber_mode) emu_barrier(); }
static inline float __shfl_sync(unsigned, float v, int src, int width = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    return emu_shfl(v, (lane / width) * width + (src % width));
}
static inline float __shfl_up_sync(unsigned, float v, unsigned d, int width = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    return emu_shfl(v, (lane % width) >= (int)d ? lane - (int)d : lane);
}
static inline float __shfl_down_sync(unsigned, float v, unsigned d, int width = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    return emu_shfl(v, (lane % width) + (int)d < width ? lane + (int)d : lane);
}
static inline float __shfl_xor_sync(unsigned, float v, int m, int width = 32) {
    const int lane = (int)(threadIdx.x & 31u);
    return emu_shfl(v, lane ^ m);
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline float __ldg(const float* p) { return *p; }
typedef void* cudaStream_t;
#define GLR_SMEM_DECL(name) float* name = emu_smem
// alignment of vector accesses and async copies: a misaligned 16-byte access is a fault on the GPU and silent on the CPU
#ifdef GLRGTV_EMU_EXACT_SMEM
#include <cstdio>
#define GLR_CHECK_ALIGN(p, n)                                                                              \
    do {                                                                                                   \
        if (((uintptr_t)(p)) & ((n) - 1)) {                                                                \
            fprintf(stderr, "glrgtv emu: %d-byte access to misaligned address %p at %s:%d\n", (int)(n), (const void*)(p), __FILE__, __LINE__); \
            abort();                                                                                       \
        }                                                                                                  \
    } while (0)
#else
#define GLR_CHECK_ALIGN(p, n) ((void)0)
#endif
// GLRGTV_EMU_EXACT_SMEM (the AddressSanitizer build, tools/emu_asan.sh): the dynamic shared memory of a launch is a heap block of
// exactly the requested size, so a kernel that indexes past its own layout trips the sanitizer instead of landing in scratch
#ifdef GLRGTV_EMU_EXACT_SMEM
struct EmuSmemScope {
    float* saved;
    void* mine;
    explicit EmuSmemScope(size_t bytes) : saved(emu_smem), mine(aligned_alloc(64, ((bytes ? bytes : 64) + 63) / 64 * 64)) {
        memset(mine, 0xFF, ((bytes ? bytes : 64) + 63) / 64 * 64);      // shared memory starts as garbage on the GPU: NaN bit patterns here
        emu_smem = (float*)mine;
    }
    ~EmuSmemScope() { emu_smem = saved; free(mine); }
};
#define GLR_EMU_SMEM_SCOPE(bytes) EmuSmemScope smem_scope_((size_t)(bytes))
#else
#define GLR_EMU_SMEM_SCOPE(bytes) ((void)0)
#endif
#define GLR_LAUNCH(kernel, grid, block, smem_bytes, stream, ...)                       \
    do {                                                                               \
        GLR_EMU_SMEM_SCOPE(smem_bytes);                                                \
        emu_async_reset_for_launch();                                                  \
        emu_dim3 g_ = (grid);                                                          \
        gridDim = g_;                                                                  \
        blockDim = emu_dim3(1, 1, 1);                                                  \
        threadIdx = emu_dim3(0, 0, 0);                                                 \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                      \
            for (unsigned by_ = 0; by_ < g_.y; ++by_)                                  \
                for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                            \
                    blockIdx = emu_dim3(bx_, by_, bz_);                                \
                    kernel(__VA_ARGS__);                                               \
                }                                                                      \
    } while (0)
// launch with real per-thread semantics: blockDim.x fibers per block (kernels that use shuffles / barriers)
#define GLR_LAUNCH_FIBERS(kernel, grid, block, smem_bytes, stream, ...)                \
    do {                                                                               \
        GLR_EMU_SMEM_SCOPE(smem_bytes);                                                \
        emu_dim3 g_ = (grid);                                                          \
        gridDim = g_;                                                                  \
        blockDim = emu_dim3((block), 1, 1);                                            \
        auto body_ = [&]() { kernel(__VA_ARGS__); };                                   \
        for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                    \
            blockIdx = emu_dim3(bx_, 0, 0);                                            \
            emu_run_block(blockDim.x, [](void* p_) { (*(decltype(body_)*)p_)(); }, &body_); \
        }                                                                              \
        blockDim = emu_dim3(1, 1, 1);                                                  \
        threadIdx = emu_dim3(0, 0, 0);                                                 \
    } while (0)
#define GLR_CHECK_LAUNCH() GLRGTV_OK
#define GLR_PROF_BEGIN(slot, stream) ((void)0)
#define GLR_PROF_END(slot, stream) ((void)0)
static inline int glr_memset_async(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
#else
// ------------------------------------------------------------------ CUDA build
#include <cuda_runtime.h>
#define GLR_CHECK_ALIGN(p, n) ((void)0)
#define GLR_SMEM_DECL(name) extern __shared__ __align__(16) float name[]
extern unsigned long long g_glr_launches;  // kernels launched by this library (bench.py reports it)
#define GLR_LAUNCH(kernel, grid, block, smem_bytes, stream, ...) \
    (++g_glr_launches, kernel<<<(grid), (block), (smem_bytes), (cudaStream_t)(stream)>>>(__VA_ARGS__))
#define GLR_LAUNCH_FIBERS GLR_LAUNCH
int glr_record_launch_error(void);
// optional per-kernel timing (glrgtv_profile_*): event pairs recorded on the launching stream
void glr_prof_mark(int slot, int end, void* stream);
extern int g_glr_prof_on;
#define GLR_PROF_BEGIN(slot, stream) do { if (g_glr_prof_on) glr_prof_mark((slot), 0, (stream)); } while (0)
#define GLR_PROF_END(slot, stream) do { if (g_glr_prof_on) glr_prof_mark((slot), 1, (stream)); } while (0)
#define GLR_CHECK_LAUNCH() glr_record_launch_error()
// Opt a kernel into `bytes` of dynamic shared memory.  The attribute is per DEVICE, so the cache is too (a process may
// drive several GPUs); `cache` is a function-local static array of GLR_MAX_DEVICES entries, one per kernel instantiation.
#define GLR_MAX_DEVICES 64
template <class K>
static inline int glr_smem_optin(K kernel, size_t bytes, size_t* cache) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return glr_record_launch_error();
    if (dev < 0 || dev >= GLR_MAX_DEVICES || bytes > cache[dev]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return glr_record_launch_error();
        if (dev >= 0 && dev < GLR_MAX_DEVICES) cache[dev] = bytes;
    }
    return 0;
}
static inline int glr_memset_async(void* p, int v, size_t n, cudaStream_t s) {
    return cudaMemsetAsync(p, v, n, s) == cudaSuccess ? 0 : -1;
}
#endif

#define GLR_THREADS 256

// ------------------------------------------------------------------ small device helpers
__host__ __device__ __forceinline__ int glr_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__host__ __device__ __forceinline__ int glr_maxi(int a, int b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ int glr_mini(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int glr_reflecti(int v, int n) {
    // torch 'reflect' padding by one pixel: -1 -> 1, n -> n-2
    if (v < 0) v = -v;
    if (v > n - 1) v = 2 * (n - 1) - v;
    return v;
}
__host__ __device__ __forceinline__ int glr_mapi(int v, int n, int pad_mode) {
    return pad_mode == GLRGTV_PAD_CLAMP ? glr_clampi(v, 0, n - 1) : glr_reflecti(v, n);
}
__host__ __device__ __forceinline__ bool glr_inside(int h, int w, int H, int W) {
    return h >= 0 && h < H && w >= 0 && w < W;
}
// all p in [0,n) with clamp(p+d,0,n-1) == q, as an inclusive range [lo,hi] (empty when lo>hi)
__host__ __device__ __forceinline__ void glr_clamp_preimage(int q, int d, int n, int& lo, int& hi) {
    lo = q - d;
    hi = q - d;
    if (q == 0) lo = 0;
    if (q == n - 1) hi = n - 1;
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
}

// five taps of the stats kernel, order c, R, D, U, L  (offsets (0,0),(0,1),(1,0),(-1,0),(0,-1))
struct StatsTaps {
    float kc, kr, kd, ku, kl;
};
__host__ __device__ __forceinline__ StatsTaps glr_taps(float p1, float pa, float pb, float p3) {
    StatsTaps t;
    t.kc = ((p1 - pa) - pb) + 4.0f * p3;
    t.kr = pa - p3;
    t.kd = pb - p3;
    t.ku = -p3;
    t.kl = -p3;
    return t;
}
__device__ __forceinline__ StatsTaps glr_load_taps(const glrgtv_stats& st, int c) {
    int i = st.n == 1 ? 0 : c;
    return glr_taps(st.p01[i], st.p02a[i], st.p02b[i], st.p03[i]);
}

// block-wide sum; result valid in thread 0 (emu: single thread).  `red` = >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
#ifdef GLRGTV_EMU
    if (!emu_fiber_mode) { (void)red; return v; }
#endif
    {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = ((int)threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
    }
}

static inline int glr_shape_ok(const glrgtv_shape* s) {
    return s && s->B > 0 && s->G > 0 && s->F > 0 && s->H > 0 && s->W > 0;
}
static inline int glr_window_ok(const glrgtv_window* w) {
    return w && w->n_edges > 0 && w->n_edges <= GLRGTV_MAX_EDGES;
}
static inline int glr_aligned(const void* p) { return p && (((uintptr_t)p) & 3u) == 0; }
#define GLR_REQUIRE_PTR(p)                               \
    do {                                                 \
        if (!glr_aligned(p)) return GLRGTV_ERR_POINTER;  \
    } while (0)

// grid helper: planes in x (limit 2^31-1), pixel chunks in y
static inline dim3 glr_grid(long planes, long pixels, int per_block) {
    long chunks = (pixels + per_block - 1) / per_block;
    if (chunks > 65535) chunks = 65535;
    return dim3((unsigned)planes, (unsigned)chunks, 1);
}
