// block_weights_bwd.cu - VJP of the edge-weight construction of one scale of the fused block, both operator
// families (GTV / GLR halves of the projected features) in one launch.  SURVEY Appendix B.1, V1X0:146-175.
//
//   fhat = f / max(|f|_F, 1e-12),  ft = M * fhat,  s_e[p] = <ft[p], ft[cl(p+d_e)]>,  w = softmax_e(s)
//
// A CTA owns one (batch, family, graph) and a TH x TW tile.  Phase 1 stages fhat, 1/|f| and the softmax VJP
//   gs_e = w_e (gw_e - sum_e' w_e' gw_e')
// on the tile (+) 1 in shared memory.  Phase 2: every similarity s_e[p] touches ft at p and at its neighbour, so
//   gft[q] = sum_e c_e ft[n_e],   c_e = gs_e[q] + gs_{opposite e}[n_e]   (+ gs_e[q] once more when n_e falls outside
// the image: the replicated neighbour IS q), then through the normalisation; gM by warp shuffles + atomics.
#include "common.cuh"

#define WB_TH 16
#define WB_TW 32
#define WB_NT 256

template <int TH, int TW>
__global__ void __launch_bounds__(WB_NT) k_block_weights_bwd(glrgtv_shape s, const float* __restrict__ feat,
                                                            const float* __restrict__ M_gtv, const float* __restrict__ M_glr,
                                                            const float* __restrict__ w_gtv, const float* __restrict__ w_glr,
                                                            const float* __restrict__ gw_gtv, const float* __restrict__ gw_glr,
                                                            float* __restrict__ gfeat, float* __restrict__ gM_gtv,
                                                            float* __restrict__ gM_glr) {
    GLR_SMEM_DECL(smem);
    const int H = s.H, W = s.W, F = s.F, G = s.G, C = G * F;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, set = (plane / G) % 2, b = plane / (2 * G);
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW;
    const size_t HW = (size_t)H * W;
    constexpr int NH = TH + 2, NW = TW + 2, NP = NH * NW;
    const size_t foff = ((size_t)b * 2 * C + (size_t)set * C + (size_t)g * F) * HW;
    const float* fp = feat + foff;
    float* gfp = gfeat + foff;
    const float* Mg = (set ? M_glr : M_gtv) + g * F;
    float* gMg = (set ? gM_glr : gM_gtv) + g * F;
    const size_t woff = ((size_t)b * G + g) * 4 * HW;
    const float* wp = (set ? w_glr : w_gtv) + woff;
    const float* gwp = (set ? gw_glr : gw_gtv) + woff;
    float* fh = smem;               // [F][NP]  fhat, clamp-extended
    float* inv = smem + F * NP;     // [NP]     1 / max(|f|, eps), negative where the norm was clamped
    float* gs = inv + NP;           // [4][NP]  softmax VJP, zero-extended
    float* red = gs + 4 * NP;       // [F]      gM partial sums
    for (int i = threadIdx.x; i < F; i += blockDim.x) red[i] = 0.f;
    // ---- phase 1
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const int hh = h0 - 1 + i / NW, ww = w0 - 1 + i % NW;
        const int h = glr_clampi(hh, 0, H - 1), w = glr_clampi(ww, 0, W - 1);
        const size_t o = (size_t)h * W + w;
        float n2 = 0.f;
        for (int f = 0; f < F; ++f) { const float v = fp[f * HW + o]; n2 += v * v; }
        const float nr = sqrtf(n2), iv = 1.f / fmaxf(nr, 1e-12f);
        for (int f = 0; f < F; ++f) fh[f * NP + i] = fp[f * HW + o] * iv;
        inv[i] = nr > 1e-12f ? iv : -iv;
        if (hh == h && ww == w) {
            float we[4], ge[4], dot = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) { we[e] = wp[e * HW + o]; ge[e] = gwp[e * HW + o]; dot += we[e] * ge[e]; }
#pragma unroll
            for (int e = 0; e < 4; ++e) gs[e * NP + i] = we[e] * (ge[e] - dot);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) gs[e * NP + i] = 0.f;
        }
    }
    __syncthreads();
    // ---- phase 2: each thread owns up to PPT pixels of the tile
    constexpr int PPT = (TH * TW + WB_NT - 1) / WB_NT;
    float c[PPT][4], dot[PPT];
    int pix[PPT];
    bool ok[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
#ifdef GLRGTV_EMU
        const int i = k;   // (emulation walks the pixels in the outer loop below)
#else
        const int i = threadIdx.x + k * WB_NT;
#endif
        pix[k] = i; ok[k] = false; dot[k] = 0.f;
        c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.f;
    }
#ifdef GLRGTV_EMU
    for (int base = 0; base < TH * TW; base += PPT) {
    for (int k = 0; k < PPT; ++k) { pix[k] = base + k; dot[k] = 0.f; }
#endif
    const int offs[4] = {-NW, -1, 1, NW}, opp[4] = {3, 2, 1, 0};
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int lh = pix[k] / TW, lw = pix[k] % TW, h = h0 + lh, w = w0 + lw;
        ok[k] = pix[k] < TH * TW && h < H && w < W;
        if (!ok[k]) continue;
        const int q = (lh + 1) * NW + (lw + 1);
        const bool in[4] = {h > 0, w > 0, w < W - 1, h < H - 1};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float own = gs[e * NP + q];
            c[k][e] = in[e] ? own + gs[opp[e] * NP + q + offs[e]] : 2.f * own;
        }
    }
    // pass 1: <fhat, M*gft> per pixel and the multiM gradient
    for (int f = 0; f < F; ++f) {
        const float m = Mg[f];
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!ok[k]) continue;
            const float* p = fh + f * NP + (pix[k] / TW + 1) * NW + (pix[k] % TW + 1);
            // ft[n] = M * fhat[n]; clamp-extended fhat makes the out-of-image neighbour equal to the pixel itself
            const float gft = m * (c[k][0] * p[-NW] + c[k][1] * p[-1] + c[k][2] * p[1] + c[k][3] * p[NW]);
            part += gft * p[0];
            dot[k] += p[0] * m * gft;
        }
#ifdef GLRGTV_EMU
        red[f] += part;
#else
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[f], part);
#endif
    }
    // pass 2: through fhat = f / max(|f|, eps)
    for (int f = 0; f < F; ++f) {
        const float m = Mg[f];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!ok[k]) continue;
            const int q = (pix[k] / TW + 1) * NW + (pix[k] % TW + 1);
            const float* p = fh + f * NP + q;
            const float gh = m * m * (c[k][0] * p[-NW] + c[k][1] * p[-1] + c[k][2] * p[1] + c[k][3] * p[NW]);
            const float iv = inv[q];
            const float v = iv > 0.f ? (gh - p[0] * dot[k]) * iv : gh * (-iv);
            gfp[f * HW + (size_t)(h0 + pix[k] / TW) * W + (w0 + pix[k] % TW)] = v;
        }
    }
#ifdef GLRGTV_EMU
    }
#endif
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) atomicAdd(&gMg[f], red[f]);
}

int glr_weights_walk_bwd(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, const float* wt, const float* wl,
                         const float* gwt, const float* gwl, float* gfeat, float* gMt, float* gMl, void* stream);

int glr_block_weights_bwd(const glrgtv_shape* s, const float* feat, const float* M_gtv, const float* M_glr,
                          const float* w_gtv, const float* w_glr, const float* gw_gtv, const float* gw_glr, float* gfeat,
                          float* gM_gtv, float* gM_glr, void* stream) {
    {   // the row walkers of weights_walk.cu where the shape is theirs, else the tile kernel below
        const int rcw = glr_weights_walk_bwd(*s, feat, M_gtv, M_glr, w_gtv, w_glr, gw_gtv, gw_glr, gfeat, gM_gtv, gM_glr, stream);
        if (rcw != GLRGTV_ERR_UNSUPPORTED) return rcw;
    }
    const long tiles = (long)((s->W + WB_TW - 1) / WB_TW) * ((s->H + WB_TH - 1) / WB_TH);
    const long blocks = tiles * s->B * 2 * s->G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    const size_t smem = ((size_t)(s->F + 5) * (WB_TH + 2) * (WB_TW + 2) + s->F + 4) * sizeof(float);
    if (smem > 200 * 1024) return GLRGTV_ERR_UNSUPPORTED;
#ifndef GLRGTV_EMU
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k_block_weights_bwd<WB_TH, WB_TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return glr_record_launch_error();
    }
#endif
    GLR_LAUNCH((k_block_weights_bwd<WB_TH, WB_TW>), dim3((unsigned)blocks), WB_NT, smem, stream, *s, feat, M_gtv, M_glr,
               w_gtv, w_glr, gw_gtv, gw_glr, gfeat, gM_gtv, gM_glr);
    return GLR_CHECK_LAUNCH();
}
