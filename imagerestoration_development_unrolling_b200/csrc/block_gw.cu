// block_gw.cu - edge-weight gradients of one backward stage of the fused block.
//
// The weight gradients of a stage need only the FIRST-level stencils of its two operands,
//     s_K = S_K z (clamp-extended),   h_K = a_K S0_K g   (g = the stage's upstream gradient),
//   L :      gw_e[p] -= sum_f  h_L[p] * s_L[n_e(p)]
//   T lin :  gw_e[p] += sum_f  2 w_e (h_T[p] - h_T[n_e]) (s_T[p] - s_T[n_e])
//   T thr :  gw_e[p] += sum_f  D phi(w d) + D w phi'(w d) d,   D = h_T[p] - h_T[n_e], d = s_T[p] - s_T[n_e]   (stage X2, upstream gB)
//            ggamma   += sum   -2 sign(w d) D w [|w d| > Gamma]   (times Gamma: the parameter is a log)
// and they are sums over the F channels of a graph.  The streaming stage kernels (block_stream_bwd.cu) give every
// channel its own walkers, so this reduction lives here: a CTA owns one (batch, graph) and one tile of one
// resolution, walks the graph's channels with the tile's operands in shared memory, and keeps the eight
// gradient values of each of its pixels in registers until the end.  The coarse resolution pools its operands
// (2x2 mean) while loading.
#include "tile.cuh"
#include "stream_bwd.cuh"

enum { GW_X3 = BW_X3, GW_X2 = BW_X2A, GW_X1 = BW_X1, GW_BA = BW_BA };   // GW_X2 covers both parts of stage X2
#define GW_TH 16
#define GW_TW 32
#define GW_NT 256
#ifndef GW_MINB
#define GW_MINB 4
#endif
#define GW_PPT ((GW_TH * GW_TW) / GW_NT)
#define GW_ZH (GW_TH + 4)
#define GW_ZW (GW_TW + 4)
#define GW_SH (GW_TH + 2)
#define GW_SW (GW_TW + 2)
#define GW_NZ ((GW_ZH * GW_ZW + GW_NT - 1) / GW_NT)     // operand-tile elements per thread
#define GW_NS ((GW_SH * GW_SW + GW_NT - 1) / GW_NT)     // stencil-plane elements per thread

// MODE GW_X2 does the linear part (upstream gA) and the thresholded part (upstream gB) of stage X2 in one pass: both read
// z = x1 and the same two source tensors.
template <int MODE, bool COARSE>
__global__ void __launch_bounds__(GW_NT, GW_MINB) k_gw_stage(GwArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool HAS_L = MODE != GW_BA, X2 = MODE == GW_X2;
    constexpr int TH = GW_TH, TW = GW_TW, ZH = GW_ZH, ZW = GW_ZW, SH = GW_SH, SW = GW_SW;
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int LH = COARSE ? H / 2 : H, LW = COARSE ? W / 2 : W;
    const int tiles_w = (LW + TW - 1) / TW, tiles_h = (LH + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW;
    const size_t HW = (size_t)H * W, LHW = (size_t)LH * LW;
    float* zt = smem;                    // [ZH][ZW] z, clamp-extended
    float* gt = zt + ZH * ZW;            // [ZH][ZW] g (X2: gA), zero-extended
    float* g2 = gt + ZH * ZW;            // [ZH][ZW] X2: gB
    float* sA = g2 + ZH * ZW;            // [SH][SW] S_L z at the clamped centre
    float* sB = sA + SH * SW;            // [SH][SW] S_T z
    float* hT = sB + SH * SW;            // [SH][SW] a_T S0_T g at the clamped centre
    float* h2 = hT + SH * SW;            // [SH][SW] X2: a_T S0_T gB
    float* red = h2 + SH * SW;           // [32]

    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const float s1 = a.p.skip ? a.p.skip[1] : 1.f, c23 = al2 * s1;
    float ca, cb = 0.f, ca2 = 0.f, cb2 = 0.f;       // g = ca src0 + cb src1 ; X2: gB = ca2 src0 + cb2 src1
    if (MODE == GW_X3) ca = -c23;
    else if (MODE == GW_X2) { ca = -be2 * c23; cb = -al1; ca2 = c23 + be2 * c23; cb2 = al1; }
    else if (MODE == GW_X1) ca = -al0;
    else ca = 1.f;
    const float aT = expf(COARSE ? a.p.ro1[g] : a.p.ro0[g]), aL = HAS_L ? expf(COARSE ? a.p.mu1[g] : a.p.mu0[g]) : 0.f;
    const float Gam = X2 ? expf(COARSE ? a.p.gamma1[g] : a.p.gamma0[g]) : 0.f;
    const glrgtv_stats& stT = COARSE ? a.p.gtv1.stats : a.p.gtv0.stats;
    const glrgtv_stats& stL = COARSE ? a.p.glr1.stats : a.p.glr0.stats;
    const float* wT = (COARSE ? a.wT1 : a.wT0) + ((size_t)b * G + g) * 4 * LHW;
    float* gwT = (COARSE ? a.gwT1 : a.gwT0) + ((size_t)b * G + g) * 4 * LHW;
    float* gwL = (COARSE ? a.gwL1 : a.gwL0) + ((size_t)b * G + g) * 4 * LHW;

    // ---- channel-independent work done once: where each thread's operand-tile and stencil-plane elements live
    int zoff[GW_NZ];      // source offset of operand-tile element i = tid + k NT (fine-resolution offset of the clamped pixel); -1: none
    bool zin[GW_NZ];      // element is inside the image (g is zero outside)
#pragma unroll
    for (int k = 0; k < GW_NZ; ++k) {
        const int i = (int)threadIdx.x + k * GW_NT;
        zoff[k] = -1; zin[k] = false;
        if (i < ZH * ZW) {
            const int hh = h0 - 2 + i / ZW, ww = w0 - 2 + i % ZW;
            const int hc = glr_clampi(hh, 0, LH - 1), wc = glr_clampi(ww, 0, LW - 1);
            zoff[k] = COARSE ? (2 * hc) * W + 2 * wc : hc * W + wc;
            zin[k] = hh == hc && ww == wc;
        }
    }
    int soff[GW_NS];      // operand-tile index of the clamped centre of stencil-plane element i; -1: none
#pragma unroll
    for (int k = 0; k < GW_NS; ++k) {
        const int i = (int)threadIdx.x + k * GW_NT;
        soff[k] = -1;
        if (i < SH * SW) {
            const int hc = glr_clampi(h0 - 1 + i / SW, 0, LH - 1), wc = glr_clampi(w0 - 1 + i % SW, 0, LW - 1);
            soff[k] = (hc - (h0 - 2)) * ZW + (wc - (w0 - 2));
        }
    }
    // this thread's pixels, their raw GTV weights (channel-independent) and gradient accumulators
    float accL[GW_PPT][4], accT[GW_PPT][4], we[GW_PPT][4], gam = 0.f;
    bool ok[GW_PPT];
    int ph[GW_PPT], pw[GW_PPT];
#pragma unroll
    for (int k = 0; k < GW_PPT; ++k) {
        const int i = (int)threadIdx.x + k * GW_NT;
        ph[k] = i / TW; pw[k] = i % TW;
        ok[k] = h0 + ph[k] < LH && w0 + pw[k] < LW;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            accL[k][e] = accT[k][e] = 0.f;
            we[k][e] = ok[k] ? wT[e * LHW + (size_t)(h0 + ph[k]) * LW + w0 + pw[k]] : 0.f;
        }
    }
    // operand at this resolution: the tensor itself, or its 2x2 mean
    auto ld = [&](const float* q) -> float { return COARSE ? 0.25f * (q[0] + q[1] + q[W] + q[W + 1]) : q[0]; };

    // operands of one channel for this thread's tile elements, fetched one channel ahead of their use
    float vz[GW_NZ], v0[GW_NZ], v1[GW_NZ];
    auto fetch = [&](int f) {
        const size_t off = ((size_t)b * G * F + (size_t)g * F + f) * HW;
#pragma unroll
        for (int k = 0; k < GW_NZ; ++k) {
            vz[k] = v0[k] = v1[k] = 0.f;
            if (zoff[k] < 0) continue;
            vz[k] = ld(a.z + off + zoff[k]);
            if (zin[k]) {
                v0[k] = ld(a.src0 + off + zoff[k]);
                if (X2) v1[k] = ld(a.src1 + off + zoff[k]);
            }
        }
    };
    fetch(0);
    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const StatsTaps kT = glr_load_taps(stT, c), kL = HAS_L ? glr_load_taps(stL, c) : kT;
#pragma unroll
        for (int k = 0; k < GW_NZ; ++k) {
            if (zoff[k] < 0) continue;
            const int i = (int)threadIdx.x + k * GW_NT;
            zt[i] = vz[k];
            gt[i] = ca * v0[k] + cb * v1[k];
            if (X2) g2[i] = ca2 * v0[k] + cb2 * v1[k];
        }
        if (f + 1 < F) fetch(f + 1);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GW_NS; ++k) {
            if (soff[k] < 0) continue;
            const int i = (int)threadIdx.x + k * GW_NT;
            const float* q = zt + soff[k];
            const float* r = gt + soff[k];
            sB[i] = kT.kc * q[0] + kT.kr * q[1] + kT.kd * q[ZW] + kT.ku * q[-ZW] + kT.kl * q[-1];
            if (HAS_L) sA[i] = kL.kc * q[0] + kL.kr * q[1] + kL.kd * q[ZW] + kL.ku * q[-ZW] + kL.kl * q[-1];
            hT[i] = aT * (kT.kc * r[0] + kT.kr * r[1] + kT.kd * r[ZW] + kT.ku * r[-ZW] + kT.kl * r[-1]);
            if (X2) {
                const float* r2 = g2 + soff[k];
                h2[i] = aT * (kT.kc * r2[0] + kT.kr * r2[1] + kT.kd * r2[ZW] + kT.ku * r2[-ZW] + kT.kl * r2[-1]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GW_PPT; ++k) {
            if (!ok[k]) continue;
            const int is = (ph[k] + 1) * SW + pw[k] + 1;
            const int offs[4] = {-SW, -1, 1, SW};
            if (HAS_L) {
                const float* r = gt + (ph[k] + 2) * ZW + pw[k] + 2;
                const float hL = aL * (kL.kc * r[0] + kL.kr * r[1] + kL.kd * r[ZW] + kL.ku * r[-ZW] + kL.kl * r[-1]);
#pragma unroll
                for (int e = 0; e < 4; ++e) accL[k][e] -= hL * sA[is + offs[e]];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float D = hT[is] - hT[is + offs[e]], d = sB[is] - sB[is + offs[e]], w = we[k][e];
                accT[k][e] += 2.f * w * D * d;
                if (X2) {
                    const float D2 = h2[is] - h2[is + offs[e]], t = w * d;
                    accT[k][e] += D2 * glr_phi(t, Gam) + D2 * w * glr_dphi(t, Gam) * d;
                    if (fabsf(t) > Gam) gam += D2 * w * (t > 0.f ? -2.f : 2.f);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < GW_PPT; ++k) {
        if (!ok[k]) continue;
        const size_t o = (size_t)(h0 + ph[k]) * LW + w0 + pw[k];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (HAS_L) gwL[e * LHW + o] = a.assign ? accL[k][e] : gwL[e * LHW + o] + accL[k][e];
            gwT[e * LHW + o] = a.assign ? accT[k][e] : gwT[e * LHW + o] + accT[k][e];
        }
    }
    if (X2) {
        const float tot = block_sum(gam, red);
        if (threadIdx.x == 0 && tot != 0.f) atomicAdd((COARSE ? a.ggamma1 : a.ggamma0) + g, tot * Gam);
    }
}

extern unsigned long long g_glr_stream_launches;
// 1 (default): the tiled kernel below; 0: the streaming form (block_gw_stream.cu) where its range allows.  Measured on
// B200 at the benchmark sizes the streaming form is level with the tiled one for X3 / X1 / BA and slower for X2 (its
// 768-thread CTA pays one block barrier per row), so the tiled kernel stays the default.
int g_glr_gw_tiled = 1;
extern "C" int glrgtv_set_gw_kernel(int streaming) { g_glr_gw_tiled = streaming ? 0 : 1; return GLRGTV_OK; }
template <int MODE>
int glr_gw_stage(const GwArgs& a, int slot, void* stream) {
    const glrgtv_shape& s = a.s;
    const size_t smem = (3 * GW_ZH * GW_ZW + 4 * GW_SH * GW_SW + 32) * sizeof(float);
    for (int lvl = 0; lvl < 2; ++lvl) {
        const int LH = lvl ? s.H / 2 : s.H, LW = lvl ? s.W / 2 : s.W;
        const long blocks = (long)((LW + GW_TW - 1) / GW_TW) * ((LH + GW_TH - 1) / GW_TH) * s.B * s.G;
        if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
        ++g_glr_stream_launches;
        GLR_PROF_BEGIN(slot, stream);
        if (lvl) GLR_LAUNCH_FIBERS((k_gw_stage<MODE, true>), dim3((unsigned)blocks), GW_NT, smem, stream, a);
        else GLR_LAUNCH_FIBERS((k_gw_stage<MODE, false>), dim3((unsigned)blocks), GW_NT, smem, stream, a);
        GLR_PROF_END(slot, stream);
        const int rc = GLR_CHECK_LAUNCH();
        if (rc) return rc;
    }
    return GLRGTV_OK;
}
template int glr_gw_stage<GW_X3>(const GwArgs&, int, void*);
template int glr_gw_stage<GW_X2>(const GwArgs&, int, void*);
template int glr_gw_stage<GW_X1>(const GwArgs&, int, void*);
template int glr_gw_stage<GW_BA>(const GwArgs&, int, void*);
