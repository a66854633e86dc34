// weights_walk.cu - edge-weight construction of one level of the fused block and its VJP as register row walkers
// (V1X0:146-175, 716-733).  Round 2 replacement of the tile kernels k_block_weights /
// k_block_weights_bwd, which ran at 83 % issue utilisation and ~0.3 of the HBM roof (profiles/r02_summary.md: ~600 / ~1200
// instructions per pixel and graph): here a pixel and graph costs ~75 / ~110.
//
//   fhat = f / max(|f|_F, 1e-12),  ft = M * fhat,  s_e[p] = <ft[p], ft[cl(p + d_e)]>,  w = softmax_e(s),  e = U, L, R, D
//
// A walker = the lanes that own one (batch, family, graph) plane's strip of up to 32 PX columns (a lane owns PX adjacent
// columns, PX = 4 or 2, vector loads) over a band of rows.  It walks top to bottom, one row per step, with the previous row in
// registers, so every feature is read from HBM once (+ 2 halo rows per band) and nothing goes through shared memory:
//   * the similarities are symmetric - s_D[r] = s_U[r+1] and s_R[c] = s_L[c+1] - so a pixel costs TWO dot products, not four;
//     the horizontal one of a lane's last column takes the neighbour lane's first column by warp shuffle; replicate padding
//     turns the out-of-image neighbour into the pixel itself (s = |ft|^2);
//   * planes wider than a strip: the first / last lane of a strip fetches the single column next door itself (scalar loads);
//   * the VJP uses the undirected-edge coefficients kH[r][c] = gs_R[r][c] + gs_L[r][c+1], kV[r][c] = gs_D[r][c] + gs_U[r+1][c]
//     (gs = softmax VJP), gft[q] = sum over the four neighbours of k * ft[n] (+ twice the border edges on the pixel itself),
//     accumulated in scatter form while walking (the row below adds its term one step later), then the VJP of the
//     normalisation; multiM gradients are per-lane sums reduced over the walker's lanes, one atomic per feature and walker.
// Shapes: W % PX == 0 and F = 6 or 12 (the shipped configurations); anything else stays on the tile kernels.
#include "common.cuh"
#include "tile.cuh"

#define WW_NT 128

#ifdef GLRGTV_EMU
__device__ __forceinline__ float ww_rsqrt(float x) { return 1.f / sqrtf(x); }
__device__ __forceinline__ float ww_exp(float x) { return expf(x); }
__device__ __forceinline__ float ww_rcp(float x) { return 1.f / x; }
#else
__device__ __forceinline__ float ww_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ float ww_exp(float x) { return __expf(x); }
__device__ __forceinline__ float ww_rcp(float x) { return __frcp_rn(x); }
#endif

template <int PX>
__device__ __forceinline__ void ww_ld(const float* p, float (&v)[PX]) {
    if constexpr (PX == 4) {
        GLR_CHECK_ALIGN(p, 16);
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        GLR_CHECK_ALIGN(p, 8);
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
}
template <int PX>
__device__ __forceinline__ void ww_st(float* p, const float (&v)[PX]) {
    if constexpr (PX == 4) {
        GLR_CHECK_ALIGN(p, 16);
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        GLR_CHECK_ALIGN(p, 8);
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
}

struct WWArgs {
    glrgtv_shape s;                 // geometry of THIS level
    const float* feat;              // [B, 2C, H, W]: channels [0,C) -> GTV family, [C,2C) -> GLR family
    const float* M_gtv; const float* M_glr;          // multiM [G, F]
    float* w_gtv; float* w_glr;                      // forward: weights out [B,G,4,H,W]
    const float* wi_gtv; const float* wi_glr;        // backward: weights in
    const float* gw_gtv; const float* gw_glr;        // backward: their gradients
    float* gfeat;                                    // backward: [B, 2C, H, W]
    float* gM_gtv; float* gM_glr;                    // backward: accumulated
    int lq, lg_lq;                  // lanes per walker (a power of two <= 32) and its log2
    int n_strips, band_rows, n_bands;
    long n_walkers;
};

struct WWGeom {
    int lane, col, g, set, b, R0;
    bool live, act, imgL, imgR, seamL, seamR;
};
template <int PX>
__device__ __forceinline__ WWGeom ww_geom(const WWArgs& a) {
    WWGeom q;
    const int tid = (int)threadIdx.x;
    q.lane = tid & (a.lq - 1);
    long wk = ((long)blockIdx.x * (long)blockDim.x + tid) >> a.lg_lq;
    q.live = wk < a.n_walkers;
    if (!q.live) wk = a.n_walkers - 1;                  // idle lanes shadow the last walker: loops and shuffles stay warp-uniform
    const int band = (int)(wk % a.n_bands); wk /= a.n_bands;
    const int strip = (int)(wk % a.n_strips); wk /= a.n_strips;
    q.g = (int)(wk % a.s.G); wk /= a.s.G;
    q.set = (int)(wk & 1); q.b = (int)(wk >> 1);
    q.col = (strip * a.lq + q.lane) * PX;
    q.act = q.col < a.s.W;
    q.imgL = q.col == 0;
    q.imgR = q.col + PX >= a.s.W;
    q.seamL = q.act && q.lane == 0 && !q.imgL;
    q.seamR = q.act && q.lane == a.lq - 1 && !q.imgR;
    q.R0 = band * a.band_rows;
    return q;
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
#ifndef WW_FWD_MINB
#define WW_FWD_MINB 4
#endif
template <int FT, int PX>
__global__ void __launch_bounds__(WW_NT, WW_FWD_MINB) k_weights_walk(WWArgs a) {
    const int H = a.s.H, W = a.s.W, G = a.s.G, C = G * FT;
    const size_t HW = (size_t)H * W;
    const WWGeom q = ww_geom<PX>(a);
    const int lq = a.lq;
    const int colc = q.act ? q.col : 0;
    const float* fp = a.feat + ((size_t)q.b * 2 * C + (size_t)q.set * C + (size_t)q.g * FT) * HW + colc;
    float* wp = (q.set ? a.w_glr : a.w_gtv) + ((size_t)q.b * G + q.g) * 4 * HW + colc;
    const float* Mg = (q.set ? a.M_glr : a.M_gtv) + q.g * FT;
    float M[FT];
#pragma unroll
    for (int f = 0; f < FT; ++f) M[f] = Mg[f];
    const bool store_ok = q.live && q.act;

    float ftp[FT][PX], sUp[PX], hp[PX], sL0p = 0.f;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        sUp[j] = hp[j] = 0.f;
#pragma unroll
        for (int f = 0; f < FT; ++f) ftp[f][j] = 0.f;
    }
    const int Rend = q.R0 + a.band_rows;             // rows [R0, Rend) (clipped to H) are this walker's
#pragma unroll 1
    for (int r = q.R0 - 1; r <= Rend; ++r) {
        const float* row = fp + (size_t)glr_clampi(r, 0, H - 1) * W;
        // ---- the row's features, normalised and scaled
        float ftc[FT][PX], n2[PX];
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            if (q.act) ww_ld<PX>(row + f * HW, ftc[f]);
            else {
#pragma unroll
                for (int j = 0; j < PX; ++j) ftc[f][j] = 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            n2[j] = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) n2[j] = fmaf(ftc[f][j], ftc[f][j], n2[j]);
            n2[j] = ww_rsqrt(fmaxf(n2[j], 1e-24f));             // 1 / max(|f|, 1e-12)
#pragma unroll
            for (int f = 0; f < FT; ++f) ftc[f][j] = ftc[f][j] * n2[j] * M[f];
        }
        // ---- horizontal similarities: h[j] = <ft[j], ft[j+1]>; the last one reaches into the next lane
        float h[PX], hr = 0.f;
#pragma unroll
        for (int j = 0; j + 1 < PX; ++j) {
            h[j] = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) h[j] = fmaf(ftc[f][j], ftc[f][j + 1], h[j]);
        }
#pragma unroll
        for (int f = 0; f < FT; ++f) hr = fmaf(ftc[f][PX - 1], __shfl_down_sync(0xffffffffu, ftc[f][0], 1, lq), hr);
        if (q.imgR) {                                  // replicate padding: the right neighbour is the pixel itself
            hr = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) hr = fmaf(ftc[f][PX - 1], ftc[f][PX - 1], hr);
        }
        if (q.seamR) {                                 // the next strip's first column
            float x[FT], m2 = 0.f, d = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) { x[f] = row[f * HW + PX]; m2 = fmaf(x[f], x[f], m2); }
#pragma unroll
            for (int f = 0; f < FT; ++f) d = fmaf(x[f] * M[f], ftc[f][PX - 1], d);
            hr = d * ww_rsqrt(fmaxf(m2, 1e-24f));
        }
        h[PX - 1] = hr;
        float sL0 = __shfl_up_sync(0xffffffffu, hr, 1, lq);
        if (q.imgL) {
            sL0 = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) sL0 = fmaf(ftc[f][0], ftc[f][0], sL0);
        }
        if (q.seamL) {                                 // the previous strip's last column
            float x[FT], m2 = 0.f, d = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) { x[f] = row[f * HW - 1]; m2 = fmaf(x[f], x[f], m2); }
#pragma unroll
            for (int f = 0; f < FT; ++f) d = fmaf(x[f] * M[f], ftc[f][0], d);
            sL0 = d * ww_rsqrt(fmaxf(m2, 1e-24f));
        }
        // ---- vertical similarity with the previous row: s_D of row r-1 = s_U of row r (clamped rows make it |ft|^2 at the border)
        float dv[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            dv[j] = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) dv[j] = fmaf(ftp[f][j], ftc[f][j], dv[j]);
        }
        // ---- row r-1 is complete: softmax over (U, L, R, D)
        const int ro = r - 1;
        if (ro >= q.R0 && ro < H && store_ok) {
            float oU[PX], oL[PX], oR[PX], oD[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const float su = sUp[j], sl = j ? hp[j ? j - 1 : 0] : sL0p, sr = hp[j], sd = dv[j];
                const float mx = fmaxf(fmaxf(su, sl), fmaxf(sr, sd));
                const float eu = ww_exp(su - mx), el = ww_exp(sl - mx), er = ww_exp(sr - mx), ed = ww_exp(sd - mx);
                const float inv = ww_rcp((eu + el) + (er + ed));
                oU[j] = eu * inv; oL[j] = el * inv; oR[j] = er * inv; oD[j] = ed * inv;
            }
            float* o = wp + (size_t)ro * W;
            ww_st<PX>(o, oU); ww_st<PX>(o + HW, oL); ww_st<PX>(o + 2 * HW, oR); ww_st<PX>(o + 3 * HW, oD);
        }
        // ---- rotate
        sL0p = sL0;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            sUp[j] = dv[j]; hp[j] = h[j];
#pragma unroll
            for (int f = 0; f < FT; ++f) ftp[f][j] = ftc[f][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
// one column of a neighbouring strip: the softmax VJP of its edge `e` (1 = L, 2 = R) and 1 / max(|f|, 1e-12)
template <int FT>
__device__ __forceinline__ float ww_seam_col(const float* f, const float* w, const float* gw, size_t HW, int e, float& iv) {
    float m2 = 0.f;
#pragma unroll
    for (int k = 0; k < FT; ++k) { const float v = f[k * HW]; m2 = fmaf(v, v, m2); }
    iv = ww_rsqrt(fmaxf(m2, 1e-24f));
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) dot = fmaf(w[k * HW], gw[k * HW], dot);
    return w[e * HW] * (gw[e * HW] - dot);
}

template <int FT, int PX>
__global__ void __launch_bounds__(WW_NT, FT > 6 ? 3 : 4) k_weights_walk_bwd(WWArgs a) {
    const int H = a.s.H, W = a.s.W, G = a.s.G, C = G * FT;
    const size_t HW = (size_t)H * W;
    const WWGeom q = ww_geom<PX>(a);
    const int lq = a.lq;
    const int colc = q.act ? q.col : 0;
    const size_t foff = ((size_t)q.b * 2 * C + (size_t)q.set * C + (size_t)q.g * FT) * HW + colc;
    const size_t woff = ((size_t)q.b * G + q.g) * 4 * HW + colc;
    const float* fp = a.feat + foff;
    float* gfp = a.gfeat + foff;
    const float* wp = (q.set ? a.wi_glr : a.wi_gtv) + woff;
    const float* gwp = (q.set ? a.gw_glr : a.gw_gtv) + woff;
    const float* Mg = (q.set ? a.M_glr : a.M_gtv) + q.g * FT;
    float pm[FT];
#pragma unroll
    for (int f = 0; f < FT; ++f) pm[f] = 0.f;
    const bool store_ok = q.live && q.act;

    float fhp[FT][PX], accp[FT][PX], invp[PX], gsDp[PX];
    bool okp[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        invp[j] = gsDp[j] = 0.f; okp[j] = false;
#pragma unroll
        for (int f = 0; f < FT; ++f) fhp[f][j] = accp[f][j] = 0.f;
    }
    const int Rend = q.R0 + a.band_rows;
#pragma unroll 1
    for (int r = q.R0 - 1; r <= Rend; ++r) {
        const bool in = r >= 0 && r < H && q.act;       // rows outside the image contribute nothing
        const size_t ro_ = (size_t)glr_clampi(r, 0, H - 1) * W;
        // ---- row r: normalised features and the softmax VJP gs_e = w_e (gw_e - sum_e' w_e' gw_e')
        float fhc[FT][PX], invc[PX], gs[4][PX];
        bool okc[PX];
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            if (in) ww_ld<PX>(fp + ro_ + f * HW, fhc[f]);
            else {
#pragma unroll
                for (int j = 0; j < PX; ++j) fhc[f][j] = 0.f;
            }
        }
        {
            float w[4][PX], gw[4][PX];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (in) { ww_ld<PX>(wp + ro_ + e * HW, w[e]); ww_ld<PX>(gwp + ro_ + e * HW, gw[e]); }
                else {
#pragma unroll
                    for (int j = 0; j < PX; ++j) w[e][j] = gw[e][j] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                float dot = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) dot = fmaf(w[e][j], gw[e][j], dot);
#pragma unroll
                for (int e = 0; e < 4; ++e) gs[e][j] = w[e][j] * (gw[e][j] - dot);
            }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            float m2 = 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) m2 = fmaf(fhc[f][j], fhc[f][j], m2);
            okc[j] = m2 > 1e-24f;
            invc[j] = ww_rsqrt(fmaxf(m2, 1e-24f));
#pragma unroll
            for (int f = 0; f < FT; ++f) fhc[f][j] *= invc[j];
        }
        // ---- vertical edge between rows r-1 and r: kV = gs_D[r-1] + gs_U[r]; row r-1 receives its last term and is complete
        float kV[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            kV[j] = (r >= 1 && r < H) ? gsDp[j] + gs[0][j] : 0.f;
#pragma unroll
            for (int f = 0; f < FT; ++f) accp[f][j] = fmaf(kV[j], fhc[f][j], accp[f][j]);
        }
        const int ro = r - 1;
        if (ro >= q.R0 && ro < H && store_ok) {
            float dot[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                dot[j] = 0.f;
#pragma unroll
                for (int f = 0; f < FT; ++f) {
                    const float m = __ldg(Mg + f);
                    pm[f] = fmaf(accp[f][j], fhp[f][j], pm[f]);
                    dot[j] = fmaf(fhp[f][j] * (m * m), accp[f][j], dot[j]);
                }
            }
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                float o[PX];
                const float m = __ldg(Mg + f);
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float gh = (m * m) * accp[f][j];
                    o[j] = okp[j] ? (gh - fhp[f][j] * dot[j]) * invp[j] : gh * invp[j];
                }
                ww_st<PX>(gfp + (size_t)ro * W + f * HW, o);
            }
        }
        // ---- horizontal edges of row r and the border edges (a replicated neighbour is the pixel itself: twice its own gs)
        float kH[PX], kHl, kS[PX];
#pragma unroll
        for (int j = 0; j + 1 < PX; ++j) kH[j] = gs[2][j] + gs[1][j + 1];
        kH[PX - 1] = gs[2][PX - 1] + __shfl_down_sync(0xffffffffu, gs[1][0], 1, lq);
        float ivl = 0.f, ivr = 0.f;                      // seam lanes: inverse norms of the column next door
        if (q.imgR) kH[PX - 1] = 0.f;
        if (q.seamR && in) kH[PX - 1] = gs[2][PX - 1] + ww_seam_col<FT>(fp + ro_ + PX, wp + ro_ + PX, gwp + ro_ + PX, HW, 1, ivr);
        kHl = __shfl_up_sync(0xffffffffu, kH[PX - 1], 1, lq);
        if (q.imgL) kHl = 0.f;
        if (q.seamL && in) kHl = gs[1][0] + ww_seam_col<FT>(fp + ro_ - 1, wp + ro_ - 1, gwp + ro_ - 1, HW, 2, ivl);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            float s = 0.f;
            if (r == 0) s += gs[0][j];
            if (r == H - 1) s += gs[3][j];
            if (j == 0 && q.imgL) s += gs[1][j];
            if (j == PX - 1 && q.imgR) s += gs[2][j];
            kS[j] = 2.f * s;
        }
        // ---- start row r's sum (left, right, self, up); rotate
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            float acc[PX];
            float rt = __shfl_down_sync(0xffffffffu, fhc[f][0], 1, lq), lt = __shfl_up_sync(0xffffffffu, fhc[f][PX - 1], 1, lq);
            if (q.seamR && in) rt = fp[ro_ + f * HW + PX] * ivr;
            if (q.seamL && in) lt = fp[ro_ + f * HW - 1] * ivl;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const float left = j ? fhc[f][j ? j - 1 : 0] : lt, right = j + 1 < PX ? fhc[f][j + 1 < PX ? j + 1 : 0] : rt;
                float v = kS[j] * fhc[f][j];
                v = fmaf(j ? kH[j ? j - 1 : 0] : kHl, left, v);
                v = fmaf(kH[j], right, v);
                acc[j] = fmaf(kV[j], fhp[f][j], v);
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) { accp[f][j] = acc[j]; fhp[f][j] = fhc[f][j]; }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) { invp[j] = invc[j]; okp[j] = okc[j]; gsDp[j] = gs[3][j]; }
    }
    // ---- multiM gradient: gM[f] = M[f] * sum_pixels acc_f * fhat_f, summed over the walker's lanes
    float* gMg = (q.set ? a.gM_glr : a.gM_gtv) + q.g * FT;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
        float v = store_ok ? pm[f] * Mg[f] : 0.f;
        for (int o = lq >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (q.lane == 0 && q.live && v != 0.f) atomicAdd(gMg + f, v);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int g_glr_weights_gen = 0;      // 0 = the walkers where the shape allows, 1 = always the round-1 tile kernels
static unsigned long long g_ww_launches = 0;
extern "C" unsigned long long glrgtv_weights_walk_launch_count(void) { return g_ww_launches; }
extern "C" int glrgtv_set_weights_kernels(int generation) {
    if (generation < 0 || generation > 1) return GLRGTV_ERR_SHAPE;
    g_glr_weights_gen = generation;
    return GLRGTV_OK;
}

template <int PX>
static bool ww_plan(const glrgtv_shape& s, WWArgs& a) {
    if (s.W % PX) return false;
    const int quads = s.W / PX;
    int lq = 1, lg = 0;
    while (lq < quads && lq < 32) { lq *= 2; ++lg; }
    a.s = s; a.lq = lq; a.lg_lq = lg;
    a.n_strips = (quads + lq - 1) / lq;
    const long planes = (long)s.B * 2 * s.G;
    // row bands: 32 rows cost 2 halo rows (6 % more reads); smaller bands only when the grid would not fill the machine
    int band = 32;
    while (band > 8 && planes * a.n_strips * ((s.H + band - 1) / band) * lq < 148L * 3 * WW_NT) band /= 2;
    if (band > s.H) band = s.H;
    a.band_rows = band;
    a.n_bands = (s.H + band - 1) / band;
    a.n_walkers = planes * a.n_strips * a.n_bands;
    return (a.n_walkers * lq + WW_NT - 1) / WW_NT <= 0x7fffffffL;
}
static bool ww_ptrs_ok(const void* const* p, int n) {
    for (int i = 0; i < n; ++i)
        if (!p[i] || !glr_aligned16(p[i])) return false;
    return true;
}

// returns GLRGTV_ERR_UNSUPPORTED when the shape is not the walkers' (the caller then takes the tile kernel)
int glr_weights_walk_fwd(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, float* wt, float* wl, void* stream) {
    if (g_glr_weights_gen == 1 || (s.F != 6 && s.F != 12)) return GLRGTV_ERR_UNSUPPORTED;
    WWArgs a = {};
    const void* ptrs[3] = {feat, wt, wl};
    const bool ok = s.F == 6 ? ww_plan<4>(s, a) : ww_plan<2>(s, a);      // F = 12: pairs keep the register windows small
    if (!ok || !ww_ptrs_ok(ptrs, 3)) return GLRGTV_ERR_UNSUPPORTED;
    a.feat = feat; a.M_gtv = Mt; a.M_glr = Ml; a.w_gtv = wt; a.w_glr = wl;
    const unsigned blocks = (unsigned)((a.n_walkers * a.lq + WW_NT - 1) / WW_NT);
    ++g_ww_launches;
    if (s.F == 6) GLR_LAUNCH_FIBERS((k_weights_walk<6, 4>), dim3(blocks), WW_NT, 0, stream, a);
    else GLR_LAUNCH_FIBERS((k_weights_walk<12, 2>), dim3(blocks), WW_NT, 0, stream, a);
    return GLR_CHECK_LAUNCH();
}

int glr_weights_walk_bwd(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, const float* wt, const float* wl,
                         const float* gwt, const float* gwl, float* gfeat, float* gMt, float* gMl, void* stream) {
    if (g_glr_weights_gen == 1 || (s.F != 6 && s.F != 12)) return GLRGTV_ERR_UNSUPPORTED;
    WWArgs a = {};
    const void* ptrs[6] = {feat, wt, wl, gwt, gwl, gfeat};
    const bool ok = ww_plan<2>(s, a);               // pairs: three rows of F values per lane must stay in registers
    if (!ok || !ww_ptrs_ok(ptrs, 6)) return GLRGTV_ERR_UNSUPPORTED;
    a.feat = feat; a.M_gtv = Mt; a.M_glr = Ml; a.wi_gtv = wt; a.wi_glr = wl; a.gw_gtv = gwt; a.gw_glr = gwl;
    a.gfeat = gfeat; a.gM_gtv = gMt; a.gM_glr = gMl;
    const unsigned blocks = (unsigned)((a.n_walkers * a.lq + WW_NT - 1) / WW_NT);
    ++g_ww_launches;
    if (s.F == 6) GLR_LAUNCH_FIBERS((k_weights_walk_bwd<6, 2>), dim3(blocks), WW_NT, 0, stream, a);
    else GLR_LAUNCH_FIBERS((k_weights_walk_bwd<12, 2>), dim3(blocks), WW_NT, 0, stream, a);
    return GLR_CHECK_LAUNCH();
}
