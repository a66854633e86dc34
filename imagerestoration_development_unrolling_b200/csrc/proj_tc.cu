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
    const float* wp;   // mode 0: the weights as smem images, [chunk][k-stage][hi | lo][NC rows x 32 floats, 128-byte swizzle] (k_proj_wprep)
    int ks;            // mode 0: k-stages = ceil(Kred / 32)
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
// 1-D bulk copy global -> shared, completion counted on an mbarrier (16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// fp32 -> TF32 hi part (round to nearest) + fp32 remainder
__device__ __forceinline__ void pt_split1(const float4 v, float4& h, float4& l) {
    uint32_t hx, hy, hz, hw;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hx) : "f"(v.x));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hy) : "f"(v.y));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hz) : "f"(v.z));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hw) : "f"(v.w));
    h.x = __uint_as_float(hx); h.y = __uint_as_float(hy); h.z = __uint_as_float(hz); h.w = __uint_as_float(hw);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
}
// raw -> hi | lo for the 16-byte pieces t, t + 128, ... < n4 of a tile (128 splitter threads; `raw` may be `hi`: in place), position
// for position so that the TMA swizzle is preserved.  Four loads in flight per thread: the first version's one-at-a-time loop left
// the splitter warps ~3/4 busy and close to being the weight-gradient kernel's limiter (profiles/r02_proj_stalls.md).
__device__ __forceinline__ void pt_split(const float4* raw, float4* hi, float4* lo, int t, int n4) {
    int i = t;
    for (; i + 384 < n4; i += 512) {
        const float4 v0 = raw[i], v1 = raw[i + 128], v2 = raw[i + 256], v3 = raw[i + 384];
        float4 h, l;
        pt_split1(v0, h, l); hi[i] = h; lo[i] = l;
        pt_split1(v1, h, l); hi[i + 128] = h; lo[i + 128] = l;
        pt_split1(v2, h, l); hi[i + 256] = h; lo[i + 256] = l;
        pt_split1(v3, h, l); hi[i + 384] = h; lo[i + 384] = l;
    }
    for (; i < n4; i += 128) {
        float4 h, l;
        pt_split1(raw[i], h, l); hi[i] = h; lo[i] = l;
    }
}
// the activation tile of a K-tail stage: four boxes of [32 k rows][32 pixels] of which only the first `krows` rows are read by
// the MMAs (a multiple of 8: whole k-steps) - the valid pieces are a prefix of EACH 4 KB box
__device__ __forceinline__ void pt_split_ktail(const float4* raw, float4* hi, float4* lo, int t, int krows) {
    const int per = krows * 8;
    for (int i = t; i < 4 * per; i += 128) {
        const int box = i / per, idx = box * 256 + (i - box * per);
        float4 h, l;
        pt_split1(raw[idx], h, l); hi[idx] = h; lo[idx] = l;
    }
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

// One accumulator tile from tensor memory to global memory (an epilogue warp: TMEM lanes 32 q4 .. 32 q4 + 31; `trow` = this
// warp's lane base + the buffer's first column).  The first version recomputed a 64-bit address, re-read kernel parameters and
// branched on two predicates for EVERY stored value - about 20 instructions per store, and the four epilogue warps turned out
// to be what bounded the kernel (profiles/r02_proj_stalls.md: 93 % of their samples inside this loop, none waiting).  Now: one
// pointer that advances by a plane per column, 32 columns per TMEM load, predicates hoisted out of the full chunks.
template <int MODE>
__device__ __forceinline__ void pt_store_tile(const ProjTcArgs& a, const PtUnit& un, uint32_t trow, int q4, int lane) {
    const int n0 = un.chunk * a.NC;
    const int ncols = a.Nout - n0 < a.NC ? a.Nout - n0 : a.NC;       // valid columns of this chunk (the rest is zero padding)
    if (MODE == 0) {
        const long p = (long)un.tile * 128 + 32 * q4 + lane;
        const bool ok = p < a.P;
        const size_t P = (size_t)a.P;
        float* col = a.out + ((size_t)un.b * a.Nout + n0) * P + (ok ? p : 0);
        int j = 0;
        for (; j + 32 <= ncols; j += 32) {
            uint32_t v[32];
            tmem_ld32(trow + j, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) __stcs(col + (size_t)i * P, __uint_as_float(v[i]));
            }
            col += 32 * P;
        }
        for (; j < ncols; j += 16) {
            uint32_t v[16];
            tmem_ld16(trow + j, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (j + i < ncols) __stcs(col + (size_t)i * P, __uint_as_float(v[i]));
            }
            col += 16 * P;
        }
    } else {
        const int m = un.tile * 128 + 32 * q4 + lane;
        const bool ok = m < a.M;
        float* dst = a.out + (size_t)(ok ? m : 0) * a.Nout + n0;
        for (int j = 0; j < ncols; j += 16) {
            uint32_t v[16];
            tmem_ld16(trow + j, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    if (j + i < ncols) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j + i), "f"(__uint_as_float(v[i])),
                                 "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                                 : "memory");
            }
        }
    }
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
    const uint32_t b_pad_ = (uint32_t)(lay.blo_off() - lay.b_off());
    const uint32_t tx_bytes = MODE == 0 ? (uint32_t)lay.a_bytes() + 2 * b_pad_ : (uint32_t)(lay.a_bytes() + lay.b_bytes());

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
                        // the weights' hi | lo tiles of this k-stage: ONE bulk copy of a ready-made shared-memory image (a tensor
                        // copy would be 2 NC row requests of 128 bytes per stage - more than the activations' 128)
                        bulk_load(st + (uint32_t)lay.b_off(), a.wp + ((size_t)un.chunk * a.ks + q) * (2 * b_pad_ / 4), 2 * b_pad_, FULL(s));
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
        for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
            const PtUnit un = pt_unit(a, u);
            for (long q = un.q0; q < un.q1; ++q, ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (uint32_t)((it / S) & 1);
                bar_wait(FULL(s), ph);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                float4* const ahi = (float4*)st;
                float4* const alo = (float4*)(st + lay.a_bytes());
                if (MODE == 0) {
                    const int left = a.Kred - (int)q * PT_STAGE_K;                    // valid k rows of this stage (whole k-steps)
                    if (left >= PT_STAGE_K) pt_split(ahi, ahi, alo, t, (int)(lay.a_bytes() / 16));
                    else pt_split_ktail(ahi, ahi, alo, t, (left + 7) & ~7);
                } else {
                    // rows of gY / channels of X beyond the matrix only feed accumulator rows / columns that are never stored
                    const int rows = a.M - un.tile * 128 < 128 ? a.M - un.tile * 128 : 128;
                    const int n0 = un.chunk * a.NC, cols = a.Nout - n0 < a.NC ? a.Nout - n0 : a.NC;
                    pt_split(ahi, ahi, alo, t, rows * 8);
                    pt_split((float4*)(st + lay.b_off()), (float4*)(st + lay.b_off()), (float4*)(st + lay.blo_off()), t, cols * 8);
                }
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
            bar_wait(ACCF(buf), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(buf * PT_ACC_COLS);
            if (un.q1 > un.q0) pt_store_tile<MODE>(a, un, trow, q4, lane);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(ACCE(buf));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Second pipeline ("landing ring", round 2).  The kernel above keeps only 16 KB of NEW HBM data in flight per 56 KB stage (the
// rest of a stage is the lo half of the split and the weights' hi | lo tiles), three stages per SM, and its splitter warps spend
// most of their time waiting for TMA data (profiles/r02_proj_stalls.md).  Here the raw fp32 tiles land in their own deep ring
// (SL slots of 16 KB, up to 8) and the splitter copies them - position for position, so the TMA swizzle is preserved - into a
// shallow ring of operand slots (SO = 2: hi | lo [| weights hi | lo]) that the MMAs read:
//   warp 0      TMA producer of the raw tiles            waits LEMPTY[l]            -> LFULL[l] (tx bytes)
//   warp 2      (activation mode) TMA producer of the pre-split weights into the operand slot   waits OEMPTY[o] -> OBFULL[o]
//   warps 8-11  splitter: LAND[l] -> OPER[o] hi | lo     waits LFULL[l], OEMPTY[o]  -> OSPLIT[o], LEMPTY[l]
//   warp 1      MMA issuer                               waits OSPLIT[o] (, OBFULL[o]) -> tcgen05.commit OEMPTY[o]; ACCF[buf] per unit
//   warps 4-7   epilogue (unchanged)
// ---------------------------------------------------------------------------------------------------
struct PtSmem2 {
    int NC, SL, SO, mode;
    __host__ __device__ size_t a_bytes() const { return 128 * PT_STAGE_K * 4; }
    __host__ __device__ size_t b_bytes() const { return (size_t)NC * PT_STAGE_K * 4; }
    __host__ __device__ size_t b_pad() const { return (b_bytes() + 1023) & ~(size_t)1023; }
    __host__ __device__ size_t oper_bytes() const { return 2 * a_bytes() + 2 * b_pad(); }
    __host__ __device__ size_t land_bytes() const { return a_bytes() + (mode == 1 ? b_pad() : 0); }
    __host__ __device__ size_t land_off() const { return (size_t)SO * oper_bytes(); }
    __host__ __device__ size_t bars_off() const { return land_off() + (size_t)SL * land_bytes(); }
    __host__ __device__ size_t total() const { return bars_off() + 512 + 1024; }
};

template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, 1) k_proj_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                           const ProjTcArgs a, const int SL, const int SO) {
    extern __shared__ uint8_t pt_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)pt_raw + 1023) & ~(uintptr_t)1023);
    PtSmem2 lay; lay.NC = a.NC; lay.SL = SL; lay.SO = SO; lay.mode = MODE;
    const uint32_t smem0 = s_addr(smem);
    const uint32_t bars = smem0 + (uint32_t)lay.bars_off();
    // barriers: LFULL[SL] | LEMPTY[SL] | OSPLIT[SO] | OBFULL[SO] | OEMPTY[SO] | accfull[2] | accempty[2] | tmem base
    auto LFULL = [&](int l) { return bars + 8u * l; };
    auto LEMPTY = [&](int l) { return bars + 8u * (SL + l); };
    auto OSPLIT = [&](int o) { return bars + 8u * (2 * SL + o); };
    auto OBFULL = [&](int o) { return bars + 8u * (2 * SL + SO + o); };
    auto OEMPTY = [&](int o) { return bars + 8u * (2 * SL + 2 * SO + o); };
    auto ACCF = [&](int b) { return bars + 8u * (2 * SL + 3 * SO + b); };
    auto ACCE = [&](int b) { return bars + 8u * (2 * SL + 3 * SO + 2 + b); };
    volatile uint32_t* tmem_slot = (volatile uint32_t*)(smem + lay.bars_off() + 8 * (2 * SL + 3 * SO + 4));
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);

    if (threadIdx.x == 0) {
        for (int l = 0; l < SL; ++l) { bar_init(LFULL(l), 1); bar_init(LEMPTY(l), 128); }
        for (int o = 0; o < SO; ++o) { bar_init(OSPLIT(o), 128); bar_init(OBFULL(o), 1); bar_init(OEMPTY(o), 1); }
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

    const uint32_t oper_bytes = (uint32_t)lay.oper_bytes(), land_bytes = (uint32_t)lay.land_bytes();
    const uint32_t land0 = smem0 + (uint32_t)lay.land_off();
    const uint32_t a_bytes = (uint32_t)lay.a_bytes(), b_pad = (uint32_t)lay.b_pad(), b_bytes = (uint32_t)lay.b_bytes();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: raw tiles into the landing ring
        if (lane == 0) {
            long it = 0;
            for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
                const PtUnit un = pt_unit(a, u);
                const int n0 = un.chunk * a.NC;
                for (long q = un.q0; q < un.q1; ++q, ++it) {
                    const int l = (int)(it % SL);
                    bar_wait(LEMPTY(l), (uint32_t)((it / SL) & 1) ^ 1u);
                    const uint32_t st = land0 + l * land_bytes;
                    bar_expect_tx(LFULL(l), MODE == 0 ? a_bytes : a_bytes + b_bytes);
                    if (MODE == 0) {
                        const int k0 = (int)q * PT_STAGE_K;
#pragma unroll
                        for (int j = 0; j < 4; ++j) tma_load_3d(st + j * 4096, &tmA, un.tile * 128 + 32 * j, k0, un.b, LFULL(l));
                    } else {
                        const int b = (int)(q / a.chunks_per_b), p0 = (int)(q % a.chunks_per_b) * PT_STAGE_K;
                        tma_load_3d(st, &tmA, p0, un.tile * 128, b, LFULL(l));
                        tma_load_3d(st + a_bytes, &tmB, p0, n0, b, LFULL(l));
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ (activation mode) the pre-split weights, per operand slot
        if (MODE == 0 && lane == 0) {
            long it = 0;
            for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
                const PtUnit un = pt_unit(a, u);
                for (long q = un.q0; q < un.q1; ++q, ++it) {
                    const int o = (int)(it % SO);
                    bar_wait(OEMPTY(o), (uint32_t)((it / SO) & 1) ^ 1u);
                    const uint32_t st = smem0 + o * oper_bytes + 2 * a_bytes;
                    bar_expect_tx(OBFULL(o), 2 * b_pad);
                    bulk_load(st, a.wp + ((size_t)un.chunk * a.ks + q) * (2 * b_pad / 4), 2 * b_pad, OBFULL(o));
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
                bar_wait(ACCE(buf), (uint32_t)((nu >> 1) & 1) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(buf * PT_ACC_COLS);
                uint32_t acc = 0;
                for (long q = un.q0; q < un.q1; ++q, ++it) {
                    const int o = (int)(it % SO);
                    const uint32_t ph = (uint32_t)((it / SO) & 1);
                    bar_wait(OSPLIT(o), ph);
                    if (MODE == 0) bar_wait(OBFULL(o), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem0 + o * oper_bytes;
                    const uint32_t ahi = st, alo = st + a_bytes, bhi = st + 2 * a_bytes, blo = bhi + b_pad;
                    int ksteps = PT_STAGE_K / 8;
                    if (MODE == 0) { const int left = (a.Kred - (int)q * PT_STAGE_K + 7) / 8; ksteps = left < ksteps ? left : ksteps; }
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t dah = MODE == 0 ? smem_desc(ahi + k * 1024, 4096, 512, 1) : smem_desc(ahi + k * 32, 16, 1024);
                        const uint64_t dal = MODE == 0 ? smem_desc(alo + k * 1024, 4096, 512, 1) : smem_desc(alo + k * 32, 16, 1024);
                        const uint64_t dbh = smem_desc(bhi + k * 32, 16, 1024), dbl = smem_desc(blo + k * 32, 16, 1024);
                        mma_tf32(d, dal, dbh, idesc, acc);
                        mma_tf32(d, dah, dbl, idesc, 1u);
                        mma_tf32(d, dah, dbh, idesc, 1u);
                        acc = 1u;
                    }
                    mma_commit(OEMPTY(o));
                }
                mma_commit(ACCF(buf));
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ splitter (128 threads): LAND[l] -> OPER[o] hi | lo
        const int t = (int)threadIdx.x - 256;
        long it = 0;
        const int na4 = (int)(a_bytes / 16);
        for (long u = blockIdx.x; u < a.units; u += gridDim.x) {
            const PtUnit un = pt_unit(a, u);
            for (long q = un.q0; q < un.q1; ++q, ++it) {
                const int l = (int)(it % SL), o = (int)(it % SO);
                bar_wait(LFULL(l), (uint32_t)((it / SL) & 1));
                bar_wait(OEMPTY(o), (uint32_t)((it / SO) & 1) ^ 1u);
                const uint8_t* src = smem + lay.land_off() + (size_t)l * land_bytes;
                uint8_t* dst = smem + (size_t)o * oper_bytes;
                if (MODE == 0) {
                    const int left = a.Kred - (int)q * PT_STAGE_K;
                    if (left >= PT_STAGE_K) pt_split((const float4*)src, (float4*)dst, (float4*)(dst + a_bytes), t, na4);
                    else pt_split_ktail((const float4*)src, (float4*)dst, (float4*)(dst + a_bytes), t, (left + 7) & ~7);
                } else {
                    const int rows = a.M - un.tile * 128 < 128 ? a.M - un.tile * 128 : 128;
                    const int n0 = un.chunk * a.NC, cols = a.Nout - n0 < a.NC ? a.Nout - n0 : a.NC;
                    pt_split((const float4*)src, (float4*)dst, (float4*)(dst + a_bytes), t, rows * 8);
                    pt_split((const float4*)(src + a_bytes), (float4*)(dst + 2 * a_bytes), (float4*)(dst + 2 * a_bytes + b_pad), t, cols * 8);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_arrive(OSPLIT(o));
                bar_arrive(LEMPTY(l));
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
            bar_wait(ACCF(buf), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(buf * PT_ACC_COLS);
            if (un.q1 > un.q0) pt_store_tile<MODE>(a, un, trow, q4, lane);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(ACCE(buf));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// weights [Nout][Kred] (or their transpose) -> shared-memory images [chunk][k-stage][hi | lo][NC rows][32 floats]: the TF32 hi part
// and the fp32 remainder of W[n0 + r][32 q + k] (zero outside the matrix) where the TMA unit would have put them in a 128-byte
// swizzled K-major tile: 16-byte chunk k / 4 of row r sits at chunk (k / 4) ^ (r & 7)
__global__ void k_proj_wprep(const float* __restrict__ W, float* __restrict__ Wp, int Nout, int Kred, int transposed, int NC, int n_chunks, int ks,
                             int b_pad_floats) {
    const long n = (long)n_chunks * ks * NC * 32;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i & 31), r = (int)((i >> 5) % NC);
        const long cq = (i >> 5) / NC;
        const int q = (int)(cq % ks), c = (int)(cq / ks);
        const int row = c * NC + r, col = q * 32 + k;
        float w = 0.f;
        if (row < Nout && col < Kred) w = transposed ? W[(size_t)col * Nout + row] : W[(size_t)row * Kred + col];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(w));
        float* tile = Wp + (size_t)cq * 2 * b_pad_floats;
        const int o = r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
        tile[o] = __uint_as_float(h);
        tile[b_pad_floats + o] = w - __uint_as_float(h);
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
int g_pt_pipeline = 0;      // 0 = the in-place stages (k_proj_tc, default), 1 = the landing-ring pipeline (k_proj_tc2) where its rings fit
template <int MODE>
int pt_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, ProjTcArgs& a, int slot, cudaStream_t st) {
    const long grid = a.units < pt_sms() ? a.units : pt_sms();
    const size_t budget = 220 * 1024 - 2048;
    if (g_pt_pipeline == 1) {
        // operand slots: two when at least three raw tiles still fit beside them, else one (wide accumulator chunks)
        PtSmem2 l2; l2.NC = a.NC; l2.mode = MODE; l2.SO = 2; l2.SL = 0;
        long SL = ((long)budget - 2 * (long)l2.oper_bytes()) / (long)l2.land_bytes();
        if (SL < 3) { l2.SO = 1; SL = ((long)budget - (long)l2.oper_bytes()) / (long)l2.land_bytes(); }
        if (SL > 8) SL = 8;
        if (SL >= 2) {
            l2.SL = (int)SL;
            static size_t optin2[GLR_MAX_DEVICES] = {0};
            if (int rc = glr_smem_optin(k_proj_tc2<MODE>, l2.total(), optin2)) return rc;
            ++g_glr_launches;
            k_proj_tc2<MODE><<<(unsigned)grid, PT_THREADS, l2.total(), st>>>(tmA, tmB, a, l2.SL, l2.SO);
            GLR_PROF_END(slot, st);
            return GLR_CHECK_LAUNCH();
        }
    }
    PtSmem lay; lay.NC = a.NC; lay.stages = 1;
    int S = (int)(budget / lay.stage_bytes());
    if (S > 6) S = 6;
    if (S < 2) return GLRGTV_ERR_UNSUPPORTED;
    a.stages = S; lay.stages = S;
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc = glr_smem_optin(k_proj_tc<MODE>, lay.total(), optin)) return rc;
    ++g_glr_launches;
    k_proj_tc<MODE><<<(unsigned)grid, PT_THREADS, lay.total(), st>>>(tmA, tmB, a);
    GLR_PROF_END(slot, st);
    return GLR_CHECK_LAUNCH();
}
}  // namespace

namespace {
// bytes of the pre-split weight images for an [nout x kred] product
size_t pt_ws_bytes(int nout, int kred) {
    if (nout <= 0 || kred <= 0) return 0;
    const int NC = pt_chunk(nout), n_chunks = (nout + NC - 1) / NC, ks = (kred + PT_STAGE_K - 1) / PT_STAGE_K;
    const size_t b_pad = ((size_t)NC * PT_STAGE_K * 4 + 1023) & ~(size_t)1023;
    return (size_t)n_chunks * ks * 2 * b_pad;
}
}  // namespace
// workspace of glrgtv_proj_gemm for a weight matrix [M, K], whichever way round it is applied
extern "C" size_t glrgtv_proj_gemm_workspace_bytes(int M, int K) {
    const size_t a = pt_ws_bytes(M, K), b = pt_ws_bytes(K, M);
    return a > b ? a : b;
}
extern "C" int glrgtv_set_proj_pipeline(int which) {
    if (which < 0 || which > 1) return GLRGTV_ERR_SHAPE;
    g_pt_pipeline = which;
    return GLRGTV_OK;
}

// transpose_w == 0:  Y[b] (M x N) = W (M x K)   . X[b] (K x N)      W [M,K], X [batch,K,N], Y [batch,M,N]
// transpose_w == 1:  Y[b] (K x N) = W^T (K x M) . X[b] (M x N)      W [M,K], X [batch,M,N], Y [batch,K,N]
extern "C" int glrgtv_proj_gemm(int transpose_w, int batch, int M, int N, int K, const float* W, const float* X, float* Y, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (N & 3)) return GLRGTV_ERR_SHAPE;
    const int nout = transpose_w ? K : M, kred = transpose_w ? M : K;
    if ((nout & 3) || (kred & 3)) return GLRGTV_ERR_UNSUPPORTED;
    if (!W || !X || !Y || !workspace || (((uintptr_t)W | (uintptr_t)X | (uintptr_t)Y | (uintptr_t)workspace) & 15u)) return GLRGTV_ERR_POINTER;
    if (workspace_bytes < pt_ws_bytes(transpose_w ? K : M, transpose_w ? M : K)) return GLRGTV_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* Wp = (float*)workspace;
    const int slot = transpose_w ? GLRGTV_SLOT_PROJ_DGRAD : GLRGTV_SLOT_PROJ_FWD;
    GLR_PROF_BEGIN(slot, st);
    ProjTcArgs a = {};
    a.mode = 0; a.Bt = batch; a.Nout = nout; a.Kred = kred; a.P = N; a.NC = pt_chunk(nout);
    a.n_chunks = (nout + a.NC - 1) / a.NC; a.tiles = (N + 127) / 128;
    a.units = (long)batch * a.tiles * a.n_chunks;
    a.out = Y;
    a.wp = Wp; a.ks = (kred + PT_STAGE_K - 1) / PT_STAGE_K;
    {
        const int b_pad_floats = (int)((((size_t)a.NC * PT_STAGE_K * 4 + 1023) & ~(size_t)1023) / 4);
        const long n = (long)a.n_chunks * a.ks * a.NC * 32;
        ++g_glr_launches;
        k_proj_wprep<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(W, Wp, nout, kred, transpose_w, a.NC, a.n_chunks, a.ks,
                                                                                                  b_pad_floats);
    }
    CUtensorMap tmA, tmB;
    if (int rc = pt_map(&tmA, X, N, kred, batch, PT_STAGE_K, true)) return rc;
    tmB = tmA;              // (activation mode reads the weights by bulk copy: the second tensor map is unused)
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
extern "C" size_t glrgtv_proj_gemm_workspace_bytes(int M, int K) { return (size_t)2 * (M > 0 ? M : 0) * (K > 0 ? K : 0) * sizeof(float) + 16; }
extern "C" int glrgtv_set_proj_pipeline(int which) { return which < 0 || which > 1 ? GLRGTV_ERR_SHAPE : GLRGTV_OK; }
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
