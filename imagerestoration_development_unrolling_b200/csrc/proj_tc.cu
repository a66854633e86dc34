// proj_tc.cu - the feature projections of MixtureGTVGLR (V1X0:556-612, 712, 725) on the 5th-generation tensor cores.
//
// patchs_features_extraction00 is a 1x1 convolution, ...01 a 2x2 stride-2 convolution (a 1x1 after space-to-depth) followed by
// a 1x1: forward and input gradient are   Y[b] (Co x P) = W (Co x Ci) . X[b] (Ci x P),   P = pixels, contiguous in X and Y;
// the weight gradient is                  gW (Co x Ci) = sum_b gY[b] (Co x P) . X[b]^T.
// They are the only GEMM-shaped work on the hot path.  The 1e-4 parity bar rules out one TF32 pass (1e-3), so every product is
// the three-pass split  a b ~ a_hi b_hi + a_lo b_hi + a_hi b_lo  (a_hi = a rounded to TF32, a_lo = a - a_hi: ~2^-21 relative),
// accumulated in fp32 in tensor memory.
//
// One persistent, warp-specialised kernel (hand-written PTX: tcgen05.mma kind::tf32, TMEM accumulators, TMA operand loads):
//   warp 0      TMA producer: one lane, cp.async.bulk.tensor into a ring of shared-memory stages (128B swizzle), mbarrier tx counts
//   warp 1      MMA issuer: one lane, 3 x (32/8) tcgen05.mma per stage into one of two TMEM accumulator buffers; tcgen05.commit
//               hands the stage back to the producer and the finished accumulator to the epilogue
//   warps 2-3   (idle: keeps the epilogue warps on warp ids 4..7 = TMEM lane quarters 0..3)
//   warps 4-7   epilogue: tcgen05.ld the accumulator (lane = pixel or weight row), store / reduce to global memory
//   warps 8-11  splitter: turn the landed fp32 tile into its TF32 hi part (in place) and lo part (second tile), then
//               fence.proxy.async so the tensor core (async proxy) sees the generic-proxy writes
// Activation mode (forward / input gradient): MMA M = 128 pixels (A = X tile, pixel-contiguous = MN-major; TF32 MN-major operands
// must use the 128B swizzle with 32-byte atoms, TMA mode SWIZZLE_128B_ATOM_32B), N = a chunk of
// the output channels, K = input channels; the weights are split once by a small prep kernel into [hi | lo] and arrive by TMA.
// Weight-gradient mode: M = 128 rows of gY, N = a chunk of X's channels, K = pixels (both operands K-major), the pixel range
// split over the grid, partial sums reduced into gW with vector red.global.add.
//
// These kernels are HBM-bound by design (K is 48..1536): X is read once and Y written once per output-channel chunk.
#include "common.cuh"

#ifndef GLRGTV_EMU
#include <cuda.h>

#define PT_STAGE_K 32                 // K elements per pipeline stage (one 128-byte swizzle row of fp32)
#define PT_THREADS 384
#define PT_ACC_COLS 256               // TMEM columns of one accumulator buffer (two buffers = the whole 512-column TMEM)

struct ProjTcArgs {
    int mode;          // 0: Y = W X (activations), 1: gW += gY X^T (weight gradient)
    int Bt;            // batch
    int Nout;          // MMA N in total: output channels (mode 0) / X channels (mode 1)
    int Kred;          // mode 0: input channels (the reduction)
    int M;             // mode 1: rows of gY
    long P;            // pixels per batch item
    int NC;            // columns of one accumulator chunk (<= 256, % 16 == 0)
    int n_chunks;      // Nout / NC
    int tiles;         // mode 0: pixel tiles per batch item; mode 1: 128-row tiles of gY
    int ksplit;        // mode 1: how many CTAs share one (row tile, chunk)
    long kchunks;      // mode 1: Bt * ceil(P / 32) reduction chunks in total
    int chunks_per_b;  // mode 1: ceil(P / 32)
    long units;
    int stages;
    float* out;
};

namespace {
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"): start address, leading / stride byte offsets (all >> 4),
// descriptor version 1 (sm_100), layout type 2 = 128-byte swizzle of 16-byte chunks, 1 = 128-byte swizzle of 32-byte chunks
// (the only layout MN-major TF32 operands may use)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}
// instruction descriptor of kind::tf32: fp32 accumulate, TF32 A and B, M = 128, N = n; a_mn: A is MN-major (pixel-contiguous)
__device__ __forceinline__ uint32_t instr_desc(int n, bool a_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

struct PtSmem {
    int NC, stages;
    __host__ __device__ size_t a_bytes() const { return 128 * PT_STAGE_K * 4; }                 // 16 KB
    __host__ __device__ size_t b_bytes() const { return (size_t)NC * PT_STAGE_K * 4; }
    __host__ __device__ size_t stage_bytes() const { return 2 * a_bytes() + 2 * ((b_bytes() + 1023) & ~(size_t)1023); }
    __host__ __device__ size_t b_off() const { return 2 * a_bytes(); }
    __host__ __device__ size_t blo_off() const { return b_off() + ((b_bytes() + 1023) & ~(size_t)1023); }
    __host__ __device__ size_t bars_off() const { return stage_bytes() * stages; }
    __host__ __device__ size_t total() const { return bars_off() + 256 + 1024; }                 // + barriers + alignment slack
};

// the reduction chunks [q0, q1) of unit `u` in weight-gradient mode, and its tile / chunk
struct PtUnit {
    int b, tile, chunk;
    long q0, q1;
};
__device__ __forceinline__ PtUnit pt_unit(const ProjTcArgs& a, long u) {
    PtUnit r;
    r.chunk = (int)(u % a.n_chunks); u /= a.n_chunks;
    if (a.mode == 0) {
        r.tile = (int)(u % a.tiles); r.b = (int)(u / a.tiles);
        r.q0 = 0; r.q1 = (a.Kred + PT_STAGE_K - 1) / PT_STAGE_K;
    } else {
        const int split = (int)(u % a.ksplit);
        r.tile = (int)(u / a.ksplit); r.b = 0;
        const long per = (a.kchunks + a.ksplit - 1) / a.ksplit;
        r.q0 = split * per;
        r.q1 = r.q0 + per < a.kchunks ? r.q0 + per : a.kchunks;
        if (r.q1 < r.q0) r.q1 = r.q0;
    }
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, 1) k_proj_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const ProjTcArgs a) {
    extern __shared__ uint8_t pt_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)pt_raw + 1023) & ~(uintptr_t)1023);      // 128B-swizzle atoms are 1024-byte aligned
    PtSmem lay; lay.NC = a.NC; lay.stages = a.stages;
    const int S = a.stages;
    const uint32_t smem0 = s_addr(smem);
    const uint32_t bars = smem0 + (uint32_t)lay.bars_off();
    // barriers: full[S] | split[S] | empty[S] | accfull[2] | accempty[2] | tmem base
    auto FULL = [&](int s) { return bars + 8u * s; };
    auto SPLIT = [&](int s) { return bars + 8u * (S + s); };
    auto EMPTY = [&](int s) { return bars + 8u * (2 * S + s); };
    auto ACCF = [&](int b) { return bars + 8u * (3 * S + b); };
    auto ACCE = [&](int b) { return bars + 8u * (3 * S + 2 + b); };
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + lay.bars_off() + 8 * (3 * S + 4));
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { bar_init(FULL(s), 1); bar_init(SPLIT(s), 128); bar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { bar_init(ACCF(b), 1); bar_init(ACCE(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr((const void*)tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t stage_bytes = (uint32_t)lay.stage_bytes();
    const uint32_t tx_bytes = MODE == 0 ? (uint32_t)(lay.a_bytes() + 2 * lay.b_bytes()) : (uint32_t)(lay.a_bytes() + lay.b_bytes());

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            long it = 0;
            for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
                const PtUnit un = pt_unit(a, u);
                const int n0 = un.chunk * a.NC;
                for (long q = un.q0; q < un.q1; ++q, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)((it / S) & 1);
                    bar_wait(EMPTY(s), ph ^ 1u);
                    const uint32_t st = smem0 + s * stage_bytes;
                    bar_expect_tx(FULL(s), tx_bytes);
                    if (MODE == 0) {
                        const int k0 = (int)q * PT_STAGE_K;
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_3d(st + j * 4096, &tmA, un.tile * 128 + 32 * j, k0, un.b, FULL(s));
                        tma_load_3d(st + (uint32_t)lay.b_off(), &tmB, k0, n0, 0, FULL(s));
                        tma_load_3d(st + (uint32_t)lay.blo_off(), &tmB, k0, n0, 1, FULL(s));
                    } else {
                        const int b = (int)(q / a.chunks_per_b), p0 = (int)(q % a.chunks_per_b) * PT_STAGE_K;
                        tma_load_3d(st, &tmA, p0, un.tile * 128, b, FULL(s));
                        tma_load_3d(st + (uint32_t)lay.b_off(), &tmB, p0, n0, b, FULL(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = instr_desc(a.NC, MODE == 0);
            long it = 0;
            int nu = 0;
            for (long u = blockIdx.x; u < a.units; u += gridDim.x, ++nu) {
                const PtUnit un = pt_unit(a, u);
                const int buf = nu & 1;
                const uint32_t aph = (uint32_t)((nu >> 1) & 1);
                bar_wait(ACCE(buf), aph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(buf * PT_ACC_COLS);
                uint32_t acc = 0;
                for (long q = un.q0; q < un.q1; ++q, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)((it / S) & 1);
                    bar_wait(SPLIT(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem0 + s * stage_bytes;
                    const uint32_t ahi = st, alo = st + (uint32_t)lay.a_bytes();
                    const uint32_t bhi = st + (uint32_t)lay.b_off(), blo = st + (uint32_t)lay.blo_off();
                    int ksteps = PT_STAGE_K / 8;
                    if (MODE == 0) { const int left = (a.Kred - (int)q * PT_STAGE_K + 7) / 8; ksteps = left < ksteps ? left : ksteps; }   // rows past Kred: TMA zero fill
                    for (int k = 0; k < ksteps; ++k) {
                        // A, MN-major (activations): 4 boxes [32 k][32 px], rows of 128 B whose 32-byte chunks are swizzled by (row & 3);
                        //   LBO = the next 32 pixels (4096 B), SBO = the next 4 k rows (512 B), k-step = the next 8 rows (1024 B).
                        // A, K-major (weight gradient): [128 rows][32 px], 16-byte chunks swizzled by (row & 7); k-step = 32 B.
                        const uint64_t dah = MODE == 0 ? smem_desc(ahi + k * 1024, 4096, 512, 1) : smem_desc(ahi + k * 32, 16, 1024);
                        const uint64_t dal = MODE == 0 ? smem_desc(alo + k * 1024, 4096, 512, 1) : smem_desc(alo + k * 32, 16, 1024);
                        const uint64_t dbh = smem_desc(bhi + k * 32, 16, 1024), dbl = smem_desc(blo + k * 32, 16, 1024);
                        mma_tf32(d, dal, dbh, idesc, acc);
                        mma_tf32(d, dah, dbl, idesc, 1u);
                        mma_tf32(d, dah, dbh, idesc, 1u);
                        acc = 1u;
                    }
                    mma_commit(EMPTY(s));
                }
                mma_commit(ACCF(buf));
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ splitter (128 threads)
        const int t = (int)threadIdx.x - 256;
        long it = 0;
        const int nb4 = MODE == 0 ? 0 : (int)(lay.b_bytes() / 16);
        for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
            const PtUnit un = pt_unit(a, u);
            for (long q = un.q0; q < un.q1; ++q, ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (uint32_t)((it / S) & 1);
                bar_wait(FULL(s), ph);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                auto split4 = [&](float4* hi, float4* lo, int n4) {
                    for (int i = t; i < n4; i += 128) {
                        const float4 v = hi[i];
                        float4 h, l;
                        uint32_t hx, hy, hz, hw;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hx) : "f"(v.x));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hy) : "f"(v.y));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hz) : "f"(v.z));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hw) : "f"(v.w));
                        h.x = __uint_as_float(hx); h.y = __uint_as_float(hy); h.z = __uint_as_float(hz); h.w = __uint_as_float(hw);
                        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                        hi[i] = h;
                        lo[i] = l;
                    }
                };
                split4((float4*)st, (float4*)(st + lay.a_bytes()), (int)(lay.a_bytes() / 16));
                if (MODE == 1) split4((float4*)(st + lay.b_off()), (float4*)(st + lay.blo_off()), nb4);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_arrive(SPLIT(s));
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (warps 4..7 = TMEM lanes 0..127)
        const int q4 = warp - 4;
        int nu = 0;
        for (long u = blockIdx.x; u < a.units; u += gridDim.x, ++nu) {
            const PtUnit un = pt_unit(a, u);
            const int buf = nu & 1;
            const uint32_t aph = (uint32_t)((nu >> 1) & 1);
            const int n0 = un.chunk * a.NC;
            const int ncols = a.Nout - n0 < a.NC ? a.Nout - n0 : a.NC;       // valid columns of this chunk (the rest is zero padding)
            bar_wait(ACCF(buf), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(buf * PT_ACC_COLS);
            if (un.q1 > un.q0) {
                if (MODE == 0) {
                    const long p = (long)un.tile * 128 + 32 * q4 + lane;
                    float* dst = a.out + ((size_t)un.b * a.Nout + n0) * a.P + p;
                    for (int j = 0; j < ncols; j += 16) {
                        uint32_t v[16];
                        tmem_ld16(trow + j, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (p < a.P) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (j + i < ncols) __stcs(dst + (size_t)(j + i) * a.P, __uint_as_float(v[i]));
                        }
                    }
                } else {
                    const int m = un.tile * 128 + 32 * q4 + lane;
                    float* dst = a.out + (size_t)m * a.Nout + n0;
                    for (int j = 0; j < ncols; j += 16) {
                        uint32_t v[16];
                        tmem_ld16(trow + j, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (m < a.M) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                if (j + i < ncols) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j + i), "f"(__uint_as_float(v[i])),
                                             "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                                             : "memory");
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(ACCE(buf));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// weights [Nout][Kred] (or their transpose) -> [2][Nout][Kred]: TF32 hi part | fp32 remainder
__global__ void k_proj_wprep(const float* __restrict__ W, float* __restrict__ Wp, int Nout, int Kred, int transposed) {
    const int n = Nout * Kred;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / Kred, k = i - r * Kred;
        const float w = transposed ? W[(size_t)k * Nout + r] : W[i];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(w));
        Wp[i] = __uint_as_float(h);
        Wp[n + i] = w - __uint_as_float(h);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled pt_encoder() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
// fp32 tensor [d2][d1][d0] (d0 contiguous), box [1][box1][32], 128-byte swizzle (of 32-byte chunks: atom32), zero fill outside
int pt_map(CUtensorMap* tm, const float* base, long d0, long d1, long d2, int box1, bool atom32 = false) {
    PFN_encodeTiled enc = pt_encoder();
    if (!enc) return GLRGTV_ERR_CUDA;
    const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    const cuuint64_t strides[2] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)box1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GLRGTV_OK : GLRGTV_ERR_CUDA;
}
int pt_chunk(int nout) {          // accumulator chunk: <= 256 columns, a multiple of 16 (columns past nout are zero padding)
    const int n = (nout + 255) / 256;
    return (((nout + n - 1) / n) + 15) & ~15;
}
int pt_sms() {
    static int sms[GLR_MAX_DEVICES] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= GLR_MAX_DEVICES) return 148;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev] > 0 ? sms[dev] : 148;
}
template <int MODE>
int pt_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, ProjTcArgs& a, int slot, cudaStream_t st) {
    PtSmem lay; lay.NC = a.NC; lay.stages = 1;
    int S = (int)((220 * 1024 - 2048) / lay.stage_bytes());
    if (S > 6) S = 6;
    if (S < 2) return GLRGTV_ERR_UNSUPPORTED;
    a.stages = S; lay.stages = S;
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc = glr_smem_optin(k_proj_tc<MODE>, lay.total(), optin)) return rc;
    const long grid = a.units < pt_sms() ? a.units : pt_sms();
    ++g_glr_launches;
    k_proj_tc<MODE><<<(unsigned)grid, PT_THREADS, lay.total(), st>>>(tmA, tmB, a);
    GLR_PROF_END(slot, st);
    return GLR_CHECK_LAUNCH();
}
}  // namespace

extern "C" size_t glrgtv_proj_gemm_workspace_bytes(int M, int K) { return (size_t)2 * (M > 0 ? M : 0) * (K > 0 ? K : 0) * sizeof(float); }

// transpose_w == 0:  Y[b] (M x N) = W (M x K)   . X[b] (K x N)      W [M,K], X [batch,K,N], Y [batch,M,N]
// transpose_w == 1:  Y[b] (K x N) = W^T (K x M) . X[b] (M x N)      W [M,K], X [batch,M,N], Y [batch,K,N]
extern "C" int glrgtv_proj_gemm(int transpose_w, int batch, int M, int N, int K, const float* W, const float* X, float* Y, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (N & 3)) return GLRGTV_ERR_SHAPE;
    const int nout = transpose_w ? K : M, kred = transpose_w ? M : K;
    if ((nout & 3) || (kred & 3)) return GLRGTV_ERR_UNSUPPORTED;
    if (!W || !X || !Y || !workspace || (((uintptr_t)W | (uintptr_t)X | (uintptr_t)Y | (uintptr_t)workspace) & 15u)) return GLRGTV_ERR_POINTER;
    if (workspace_bytes < glrgtv_proj_gemm_workspace_bytes(M, K)) return GLRGTV_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* Wp = (float*)workspace;
    const int slot = transpose_w ? GLRGTV_SLOT_PROJ_DGRAD : GLRGTV_SLOT_PROJ_FWD;
    GLR_PROF_BEGIN(slot, st);
    ++g_glr_launches;
    k_proj_wprep<<<(nout * kred + 255) / 256 < 1024 ? (nout * kred + 255) / 256 : 1024, 256, 0, st>>>(W, Wp, nout, kred, transpose_w);
    ProjTcArgs a = {};
    a.mode = 0; a.Bt = batch; a.Nout = nout; a.Kred = kred; a.P = N; a.NC = pt_chunk(nout);
    a.n_chunks = (nout + a.NC - 1) / a.NC; a.tiles = (N + 127) / 128;
    a.units = (long)batch * a.tiles * a.n_chunks;
    a.out = Y;
    CUtensorMap tmA, tmB;
    if (int rc = pt_map(&tmA, X, N, kred, batch, PT_STAGE_K, true)) return rc;
    if (int rc = pt_map(&tmB, Wp, kred, nout, 2, a.NC)) return rc;
    return pt_launch<0>(tmA, tmB, a, slot, st);
}

// gW [M,K] += sum_b gY[b] (M x N) . X[b]^T (N x K)     (ACCUMULATES: the caller zeroes gW)
extern "C" int glrgtv_proj_wgrad(int batch, int M, int N, int K, const float* gY, const float* X, float* gW, void* stream) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (N & 3)) return GLRGTV_ERR_SHAPE;
    if ((K & 3) || (M & 3)) return GLRGTV_ERR_UNSUPPORTED;
    if (!gY || !X || !gW || (((uintptr_t)gY | (uintptr_t)X | (uintptr_t)gW) & 15u)) return GLRGTV_ERR_POINTER;
    ProjTcArgs a = {};
    a.mode = 1; a.Bt = batch; a.Nout = K; a.M = M; a.P = N; a.NC = pt_chunk(K);
    a.n_chunks = (K + a.NC - 1) / a.NC; a.tiles = (M + 127) / 128;
    a.chunks_per_b = (N + PT_STAGE_K - 1) / PT_STAGE_K;
    a.kchunks = (long)batch * a.chunks_per_b;
    const int base = a.tiles * a.n_chunks;
    int ks = (2 * pt_sms() + base - 1) / base;                     // two units per SM: the epilogue of one overlaps the next
    if (ks > a.kchunks / 8) ks = (int)(a.kchunks / 8);
    if (ks < 1) ks = 1;
    a.ksplit = ks;
    a.units = (long)base * ks;
    a.out = gW;
    CUtensorMap tmA, tmB;
    if (int rc = pt_map(&tmA, gY, N, M, batch, 128)) return rc;
    if (int rc = pt_map(&tmB, X, N, K, batch, a.NC)) return rc;
    GLR_PROF_BEGIN(GLRGTV_SLOT_PROJ_WGRAD, stream);
    return pt_launch<1>(tmA, tmB, a, GLRGTV_SLOT_PROJ_WGRAD, (cudaStream_t)stream);
}

#else
// ------------------------------------------------------------------ CPU emulation build (tests only): plain loops
extern "C" size_t glrgtv_proj_gemm_workspace_bytes(int M, int K) { return (size_t)2 * (M > 0 ? M : 0) * (K > 0 ? K : 0) * sizeof(float); }
extern "C" int glrgtv_proj_gemm(int transpose_w, int batch, int M, int N, int K, const float* W, const float* X, float* Y, void*, size_t,
                                void*) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (N & 3)) return GLRGTV_ERR_SHAPE;
    const int nout = transpose_w ? K : M, kred = transpose_w ? M : K;
    if ((nout & 3) || (kred & 3)) return GLRGTV_ERR_UNSUPPORTED;
    for (int b = 0; b < batch; ++b)
        for (int o = 0; o < nout; ++o)
            for (int p = 0; p < N; ++p) {
                double acc = 0;
                for (int k = 0; k < kred; ++k) acc += (double)(transpose_w ? W[(size_t)k * K + o] : W[(size_t)o * K + k]) * X[((size_t)b * kred + k) * N + p];
                Y[((size_t)b * nout + o) * N + p] = (float)acc;
            }
    return GLRGTV_OK;
}
extern "C" int glrgtv_proj_wgrad(int batch, int M, int N, int K, const float* gY, const float* X, float* gW, void*) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (N & 3)) return GLRGTV_ERR_SHAPE;
    if ((K & 3) || (M & 3)) return GLRGTV_ERR_UNSUPPORTED;
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            double acc = 0;
            for (int b = 0; b < batch; ++b)
                for (int p = 0; p < N; ++p) acc += (double)gY[((size_t)b * M + m) * N + p] * X[((size_t)b * K + k) * N + p];
            gW[(size_t)m * K + k] += (float)acc;
        }
    return GLRGTV_OK;
}
#endif
