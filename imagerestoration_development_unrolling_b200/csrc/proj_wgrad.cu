// proj_wgrad.cu - weight gradient of the feature projections (1x1 / space-to-depth 2x2 convolutions, V1X0:556-612):
//     gW[m,k] += sum_b sum_n gY[b,m,n] * X[b,k,n]          gY [B,M,N], X [B,K,N] row-major, N = pixels
// A GEMM with a tiny output (M, K <= a few hundred) and a reduction over B*N ~ 2M elements.  cuBLAS' batched GEMM + sum
// leaves most SMs idle here (1.65 ms for the scale-0 shape, tools/proj_times.py); this kernel splits the REDUCTION
// across the grid: a CTA owns one 96x48 output tile and a contiguous range of 32-pixel steps, streams both operands
// through a double-buffered cp.async pipeline, keeps a 6x3 register tile per thread (fp32 FMA, exact fp32 products),
// and adds its partial tile to gW with one atomic per value at the end.
#include "tile.cuh"

#define WG_TM 96
#define WG_TK 48
#define WG_KC 32          // pixels per step
#define WG_NT 256
#define WG_P (WG_KC + 4)  // shared-memory row pitch (floats): rows stay 16-byte aligned, consecutive rows shift banks

// swap == 0: A = gY (rows m), Bm = X (rows k): tile (m0 + 96, k0 + 48) of gW [M,K]
// swap == 1: A = X (rows k), Bm = gY (rows m): tile (k0 + 96, m0 + 48), written transposed into gW [M,K]
__global__ void __launch_bounds__(WG_NT) k_proj_wgrad(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ gW,
                                                     int batch, int RA, int RB, int N, int tiles_a, int tiles_b, int splits, int swap,
                                                     int ldw) {
    GLR_SMEM_DECL(smem);
    const int tid = (int)threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tile = (int)blockIdx.x % (tiles_a * tiles_b), split = (int)blockIdx.x / (tiles_a * tiles_b);
    const int a0 = (tile / tiles_b) * WG_TM, b0 = (tile % tiles_b) * WG_TK;
    const int steps_per_b = N / WG_KC;
    const long total = (long)batch * steps_per_b;
    const long s_lo = total * split / splits, s_hi = total * (split + 1) / splits;
    float* As[2] = {smem, smem + (WG_TM + WG_TK) * WG_P};
    float* Bs[2] = {As[0] + WG_TM * WG_P, As[1] + WG_TM * WG_P};

    auto stage = [&](long s, int buf) {
        const int b = (int)(s / steps_per_b), n0 = (int)(s % steps_per_b) * WG_KC;
        // (96 + 48) rows x 8 quads = 1152 16-byte copies, 4.5 per thread
        for (int i = tid; i < (WG_TM + WG_TK) * (WG_KC / 4); i += WG_NT) {
            const int row = i / (WG_KC / 4), q = i % (WG_KC / 4);
            if (row < WG_TM) cp_async16(As[buf] + row * WG_P + 4 * q, A + ((size_t)b * RA + a0 + row) * N + n0 + 4 * q);
            else cp_async16(Bs[buf] + (row - WG_TM) * WG_P + 4 * q, Bm + ((size_t)b * RB + b0 + row - WG_TM) * N + n0 + 4 * q);
        }
    };
    float acc[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;

    if (s_lo < s_hi) { stage(s_lo, 0); }
    cp_async_commit();
    for (long s = s_lo; s < s_hi; ++s) {
        const int buf = (int)((s - s_lo) & 1);
        if (s + 1 < s_hi) stage(s + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait_pending<1>();
        __syncthreads();
        const float* as = As[buf] + (ty * 6) * WG_P;
        const float* bs = Bs[buf] + (tx * 3) * WG_P;
#pragma unroll
        for (int kk = 0; kk < WG_KC; kk += 4) {
            float a[6][4], bv[3][4];
#pragma unroll
            for (int i = 0; i < 6; ++i) ld4(as + i * WG_P + kk, a[i]);
#pragma unroll
            for (int j = 0; j < 3; ++j) ld4(bs + j * WG_P + kk, bv[j]);
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[i][j] += a[i][u] * bv[j][u];
        }
        __syncthreads();
    }
    cp_async_wait_all();
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int ra = a0 + ty * 6 + i, rb = b0 + tx * 3 + j;
            if (acc[i][j] != 0.f) atomicAdd(swap ? gW + (size_t)rb * ldw + ra : gW + (size_t)ra * ldw + rb, acc[i][j]);
        }
}

// gW [M,K] (row-major) += sum_b gY[b] (M x N) . X[b]^T (N x K).  Needs N % 32 == 0 and (M % 96 == 0, K % 48 == 0) or
// (K % 96 == 0, M % 48 == 0); GLRGTV_ERR_UNSUPPORTED otherwise (the caller then uses a library GEMM).
extern "C" int glrgtv_proj_wgrad(int batch, int M, int N, int K, const float* gY, const float* X, float* gW, void* stream) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return GLRGTV_ERR_SHAPE;
    if (!gY || !X || !gW || (((uintptr_t)gY | (uintptr_t)X) & 15u) || (((uintptr_t)gW) & 3u)) return GLRGTV_ERR_POINTER;
    if (N % WG_KC) return GLRGTV_ERR_UNSUPPORTED;
    int swap;
    if (M % WG_TM == 0 && K % WG_TK == 0) swap = 0;
    else if (K % WG_TM == 0 && M % WG_TK == 0) swap = 1;
    else return GLRGTV_ERR_UNSUPPORTED;
    const int RA = swap ? K : M, RB = swap ? M : K;
    const int tiles_a = RA / WG_TM, tiles_b = RB / WG_TK, tiles = tiles_a * tiles_b;
    const long total = (long)batch * (N / WG_KC);
    // enough CTAs for ~4 waves of 148 SMs x 3 resident CTAs, at least 8 steps each
    long splits = (148L * 3 * 4 + tiles - 1) / tiles;
    if (splits > total / 8) splits = total / 8;
    if (splits < 1) splits = 1;
    const size_t smem = (size_t)2 * (WG_TM + WG_TK) * WG_P * sizeof(float);
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_proj_wgrad, smem, optin)) return rc_;
#endif
    GLR_LAUNCH_FIBERS(k_proj_wgrad, dim3((unsigned)(tiles * splits)), WG_NT, smem, stream, swap ? X : gY, swap ? gY : X, gW, batch, RA, RB,
                      N, tiles_a, tiles_b, (int)splits, swap, K);
    return GLR_CHECK_LAUNCH();
}
