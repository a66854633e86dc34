// api_misc.cu - version / device / error plumbing of the C ABI.
#include "common.cuh"

#ifdef GLRGTV_EMU
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
alignas(64) static thread_local float emu_smem_storage[64 * 1024];
thread_local float* emu_smem = emu_smem_storage;

// ---- cooperative fibers: one per CUDA thread of the block being emulated (ucontext, round-robin) ----
#include <ucontext.h>
#include <vector>
thread_local bool emu_fiber_mode = false;
namespace {
constexpr size_t kStack = 256 * 1024;
struct FiberRt {
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    ucontext_t sched;
    unsigned n = 0, cur = 0, alive = 0;
    unsigned bar_count = 0, bar_gen = 0;
    std::vector<float> buf[2];
    std::vector<unsigned> cnt, gen;
    void (*fn)(void*) = nullptr;
    void* arg = nullptr;
};
thread_local FiberRt rt;
void fiber_yield() { swapcontext(&rt.ctx[rt.cur], &rt.sched); }
void fiber_main() {
    rt.fn(rt.arg);
    rt.done[rt.cur] = 1;
    swapcontext(&rt.ctx[rt.cur], &rt.sched);
}
}  // namespace
// ---- late-completing asynchronous copies (GLRGTV_EMU_ASYNC=late), see common.cuh ----
bool emu_async_late = false;
namespace {
struct PendingCopy { float* dst; const float* src; int n; unsigned group; };
struct PendingBulk { float* dst; const float* src; unsigned bytes; const void* bar; };
struct AsyncRt {
    std::vector<std::vector<PendingCopy>> q;      // per CUDA thread of the block
    std::vector<unsigned> open_group;             // id of the group the thread's next copies belong to (= groups committed so far)
    std::vector<PendingBulk> bulk;                // per block
};
thread_local AsyncRt art;
unsigned async_slot() { return emu_fiber_mode ? rt.cur : 0u; }
void async_reset(unsigned nthreads) {
    const char* e = getenv("GLRGTV_EMU_ASYNC");
    emu_async_late = e && !strcmp(e, "late");
    art.q.assign(nthreads ? nthreads : 1, {});
    art.open_group.assign(nthreads ? nthreads : 1, 0u);
    art.bulk.clear();
}
}  // namespace
void emu_async_reset_for_launch() { async_reset(1); }
void emu_async_push(float* dst, const float* src, int n) {
    const unsigned t = async_slot();
    art.q[t].push_back({dst, src, n, art.open_group[t]});
}
void emu_async_commit() { ++art.open_group[async_slot()]; }
void emu_async_wait(int keep) {
    const unsigned t = async_slot();
    const unsigned committed = art.open_group[t];                      // groups 0 .. committed-1 are closed
    const unsigned done_below = committed > (unsigned)keep ? committed - (unsigned)keep : 0u;
    auto& q = art.q[t];
    size_t w = 0;
    for (size_t i = 0; i < q.size(); ++i) {
        if (q[i].group < done_below) { for (int j = 0; j < q[i].n; ++j) q[i].dst[j] = q[i].src[j]; }
        else q[w++] = q[i];
    }
    q.resize(w);
}
void emu_bulk_push(float* dst, const float* src, unsigned bytes, const void* bar) { art.bulk.push_back({dst, src, bytes, bar}); }
void emu_bulk_wait(const void* bar) {
    size_t w = 0;
    for (size_t i = 0; i < art.bulk.size(); ++i) {
        if (art.bulk[i].bar == bar) memcpy(art.bulk[i].dst, art.bulk[i].src, art.bulk[i].bytes);
        else art.bulk[w++] = art.bulk[i];
    }
    art.bulk.resize(w);
}
void emu_barrier() {
    const unsigned g = rt.bar_gen;
    if (++rt.bar_count >= rt.alive) { rt.bar_count = 0; ++rt.bar_gen; return; }
    while (rt.bar_gen == g) fiber_yield();
}
float emu_shfl(float v, int src_lane) {
    if (!emu_fiber_mode) return v;
    const unsigned me = rt.cur, w = me >> 5, base = w << 5;
    const unsigned wsize = rt.n - base < 32u ? rt.n - base : 32u;
    const unsigned g = rt.gen[w];
    rt.buf[g & 1][me] = v;
    if (++rt.cnt[w] == wsize) { rt.cnt[w] = 0; ++rt.gen[w]; }
    else while (rt.gen[w] == g) fiber_yield();
    return rt.buf[g & 1][base + (unsigned)src_lane];
}
void emu_run_block(unsigned nthreads, void (*fn)(void*), void* arg) {
    rt.n = nthreads; rt.alive = nthreads; rt.fn = fn; rt.arg = arg;
    rt.bar_count = 0; rt.bar_gen = 0;
    async_reset(nthreads);
    if (rt.ctx.size() < nthreads) {
        const size_t old = rt.ctx.size();
        rt.ctx.resize(nthreads);
        rt.stacks.resize(nthreads, nullptr);
        for (size_t i = old; i < nthreads; ++i) rt.stacks[i] = (char*)malloc(kStack);
    }
    rt.done.assign(nthreads, 0);
    rt.buf[0].assign(nthreads, 0.f); rt.buf[1].assign(nthreads, 0.f);
    rt.cnt.assign((nthreads + 31) / 32, 0); rt.gen.assign((nthreads + 31) / 32, 0);
    for (unsigned i = 0; i < nthreads; ++i) {
        getcontext(&rt.ctx[i]);
        rt.ctx[i].uc_stack.ss_sp = rt.stacks[i];
        rt.ctx[i].uc_stack.ss_size = kStack;
        rt.ctx[i].uc_link = &rt.sched;
        makecontext(&rt.ctx[i], fiber_main, 0);
    }
    emu_fiber_mode = true;
    unsigned remaining = nthreads;
    // GLRGTV_EMU_SCHED = reverse | random[:seed]: the order in which the fibers of a block run between two synchronisation points.
    // A kernel whose shared-memory hand-overs are all separated by a barrier (or a warp shuffle) gives the same result under every
    // order; one that relies on "the lower thread ran first" does not - a race detector for the barrier discipline (tools/emu_races.sh)
    const char* sched = getenv("GLRGTV_EMU_SCHED");
    const int mode = !sched ? 0 : !strncmp(sched, "reverse", 7) ? 1 : !strncmp(sched, "random", 6) ? 2 : 0;
    unsigned long long lcg = 0x9E3779B97F4A7C15ull ^ (mode == 2 && sched[6] == ':' ? strtoull(sched + 7, nullptr, 10) : 0ull);
    std::vector<unsigned> order(nthreads);
    for (unsigned i = 0; i < nthreads; ++i) order[i] = mode == 1 ? nthreads - 1 - i : i;
    while (remaining) {
        if (mode == 2)
            for (unsigned i = nthreads; i > 1; --i) {                // a fresh permutation for every sweep
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[i - 1], order[(unsigned)((lcg >> 33) % i)]);
            }
        for (unsigned k = 0; k < nthreads; ++k) {
            const unsigned i = order[k];
            if (rt.done[i]) continue;
            rt.cur = i;
            threadIdx = emu_dim3(i, 0, 0);
            swapcontext(&rt.sched, &rt.ctx[i]);
            if (rt.done[i]) {
                --remaining; --rt.alive;
                if (rt.bar_count && rt.bar_count >= rt.alive) { rt.bar_count = 0; ++rt.bar_gen; }
            }
        }
    }
    emu_fiber_mode = false;
}
extern "C" {
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return "emulation build"; }
int glrgtv_check_device(void) { return GLRGTV_OK; }
unsigned long long glrgtv_launch_count(void) { return 0; }
int glrgtv_profile_enable(int) { return 0; }
int glrgtv_profile_read(float*, int*, int) { return 0; }
}
#else
#include <string.h>
static thread_local char g_last_error[256] = "";

int glr_record_launch_error(void) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return GLRGTV_OK;
    strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
    return GLRGTV_ERR_CUDA;
}

unsigned long long g_glr_launches = 0;
int g_glr_prof_on = 0;
#define GLR_PROF_SLOTS 32
#define GLR_PROF_POOL 8192
static cudaEvent_t g_ev[GLR_PROF_POOL][2];
static int g_ev_slot[GLR_PROF_POOL];
static int g_ev_made = 0, g_ev_used = 0, g_ev_open[GLR_PROF_SLOTS];

void glr_prof_mark(int slot, int end, void* stream) {
    if (slot < 0 || slot >= GLR_PROF_SLOTS) return;
    if (!end) {
        if (g_ev_used >= GLR_PROF_POOL) { g_ev_open[slot] = -1; return; }
        if (g_ev_used >= g_ev_made) {
            if (cudaEventCreate(&g_ev[g_ev_made][0]) != cudaSuccess || cudaEventCreate(&g_ev[g_ev_made][1]) != cudaSuccess) {
                g_ev_open[slot] = -1;
                return;
            }
            ++g_ev_made;
        }
        g_ev_slot[g_ev_used] = slot;
        g_ev_open[slot] = g_ev_used;
        cudaEventRecord(g_ev[g_ev_used][0], (cudaStream_t)stream);
        ++g_ev_used;
    } else if (g_ev_open[slot] >= 0) {
        cudaEventRecord(g_ev[g_ev_open[slot]][1], (cudaStream_t)stream);
        g_ev_open[slot] = -1;
    }
}

extern "C" {
unsigned long long glrgtv_launch_count(void) { return g_glr_launches; }
// start (on=1: also clears) / stop recording one CUDA-event pair around every kernel of the block entry points
int glrgtv_profile_enable(int on) {
    if (on) {
        g_ev_used = 0;
        for (int i = 0; i < GLR_PROF_SLOTS; ++i) g_ev_open[i] = -1;
    }
    g_glr_prof_on = on;
    return GLRGTV_OK;
}
// synchronises on the recorded events; ms[slot] = summed duration, count[slot] = launches.  returns #pairs
int glrgtv_profile_read(float* ms, int* count, int n_slots) {
    for (int i = 0; i < n_slots; ++i) { ms[i] = 0.f; count[i] = 0; }
    int n = 0;
    for (int i = 0; i < g_ev_used; ++i) {
        float t = 0.f;
        if (cudaEventSynchronize(g_ev[i][1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t, g_ev[i][0], g_ev[i][1]) != cudaSuccess) continue;
        if (g_ev_slot[i] < n_slots) { ms[g_ev_slot[i]] += t; count[g_ev_slot[i]] += 1; ++n; }
    }
    return n;
}
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return g_last_error; }
int glrgtv_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    return major == 10 ? GLRGTV_OK : GLRGTV_ERR_DEVICE;
}
}
#endif

#ifndef GLRGTV_EMU
// ---- space-to-depth / depth-to-space of the 2x2 stride-2 projection (patchs_features_extraction01[0], V1X0:593-603):
// out[b, c*4 + dy*2 + dx, h', w'] = x[b, c, 2h'+dy, 2w'+dx]  (torch pixel_unshuffle order).  torch's permute-copy kernel
// is uncoalesced (0.4 ms for the scale-0 map); here a thread moves a 2 x 8 block with float4 loads and stores.
__global__ void __launch_bounds__(256) k_space_to_depth(const float* __restrict__ x, float* __restrict__ y, long planes, int H, int W,
                                                       int inverse) {
    const int Hc = H / 2, Wc = W / 2, Q = Wc / 4;                    // Q: groups of 4 coarse (8 fine) columns per row
    const long total = planes * Hc * Q;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % Q), hc = (int)((i / Q) % Hc);
        const long pl = i / ((long)Q * Hc);
        float* fine = const_cast<float*>(inverse ? y : x) + (pl * H + 2 * hc) * W + 8 * q;          // the full-resolution tensor
        float* deep = const_cast<float*>(inverse ? x : y) + ((pl * 4) * Hc + hc) * (long)Wc + 4 * q;  // plane c*4 of the deep tensor
        const long ps = (long)Hc * Wc;
        if (!inverse) {
            const float4 a0 = *reinterpret_cast<const float4*>(fine), a1 = *reinterpret_cast<const float4*>(fine + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(fine + W), b1 = *reinterpret_cast<const float4*>(fine + W + 4);
            *reinterpret_cast<float4*>(deep) = make_float4(a0.x, a0.z, a1.x, a1.z);
            *reinterpret_cast<float4*>(deep + ps) = make_float4(a0.y, a0.w, a1.y, a1.w);
            *reinterpret_cast<float4*>(deep + 2 * ps) = make_float4(b0.x, b0.z, b1.x, b1.z);
            *reinterpret_cast<float4*>(deep + 3 * ps) = make_float4(b0.y, b0.w, b1.y, b1.w);
        } else {
            const float4 d0 = *reinterpret_cast<const float4*>(deep), d1 = *reinterpret_cast<const float4*>(deep + ps);
            const float4 d2 = *reinterpret_cast<const float4*>(deep + 2 * ps), d3 = *reinterpret_cast<const float4*>(deep + 3 * ps);
            *reinterpret_cast<float4*>(fine) = make_float4(d0.x, d1.x, d0.y, d1.y);
            *reinterpret_cast<float4*>(fine + 4) = make_float4(d0.z, d1.z, d0.w, d1.w);
            *reinterpret_cast<float4*>(fine + W) = make_float4(d2.x, d3.x, d2.y, d3.y);
            *reinterpret_cast<float4*>(fine + W + 4) = make_float4(d2.z, d3.z, d2.w, d3.w);
        }
    }
}

// inverse == 0: x [planes,H,W] -> y [planes*4,H/2,W/2];  inverse == 1: x [planes*4,H/2,W/2] -> y [planes,H,W].  W % 8 == 0, H even.
extern "C" int glrgtv_space_to_depth(int inverse, long planes, int H, int W, const float* x, float* y, void* stream) {
    if (planes <= 0 || H <= 0 || W <= 0 || (H & 1)) return GLRGTV_ERR_SHAPE;
    if (W % 8) return GLRGTV_ERR_UNSUPPORTED;
    if (!x || !y || (((uintptr_t)x | (uintptr_t)y) & 15u)) return GLRGTV_ERR_POINTER;
    const long total = planes * (H / 2) * (W / 8);
    const long blocks = (total + 255) / 256;
    GLR_PROF_BEGIN(inverse ? GLRGTV_SLOT_PROJ_DGRAD : GLRGTV_SLOT_PROJ_FWD, stream);
    GLR_LAUNCH(k_space_to_depth, dim3((unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks)), 256, 0, stream, x, y, planes, H, W, inverse);
    GLR_PROF_END(inverse ? GLRGTV_SLOT_PROJ_DGRAD : GLRGTV_SLOT_PROJ_FWD, stream);
    return GLR_CHECK_LAUNCH();
}
#endif
