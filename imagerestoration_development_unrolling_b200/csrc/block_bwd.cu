// block_bwd.cu - fused backward of LocalLowpassFilteringBlock / MixtureGTVGLR.
//
// The reverse sweep mirrors the four forward stages.  Each kernel re-reads the saved stage input z
// (x2, x1, bA, y), RECOMPUTES the forward operator chain in shared memory (edge signals are never stored),
// runs the adjoint chain on the incoming gradient, and emits
//   - the gradient wrt z                                  (one tensor write),
//   - this stage's contribution to the four edge-weight gradients (accumulated over the graph's F
//     channels in shared memory, then one read-modify-write per tile),
//   - per-channel stats_kernel_p* gradients and per-graph scalar gradients (block reduction + atomics).
//
//   BWD_X3:  gout, x2, bB, r1, x      -> gx2                 (r2 = bB - A x2, u2, x3 recomputed)
//   BWD_X2:  gout, gx2, x1, r1        -> gx1                 (through r1 = bB - A x1 and bB = y + R_thr x1)
//   BWD_X1:  gx1, bA                  -> gbA                 (x1 = (1+a0) bA - a0 A bA)
//   BWD_BA:  gbA, gout, gx2, y        -> gx (direct path)    (bA = y + R_lin y; + d bB/dy + skip)
//   then k_edge_weights_bwd_* per weight set -> gfeat0 / gfeat1, gmultiM.
//
// Upstream gradients that are pointwise functions of (gout, gx2) are rebuilt on the fly instead of being
// stored:  gr2 = a2 s1 gout,  gr1 = b2 gr2 + a1 gx2,  gbB = gr2 + gr1.
#include "tile.cuh"

enum { BWD_X3 = 0, BWD_X2 = 1, BWD_X1 = 2, BWD_BA = 3 };

struct BlockBwdArgs {
    glrgtv_shape s;
    glrgtv_block_params p;
    glrgtv_block_grads gr;
    const float* z;     // x2 | x1 | bA | y
    const float* gout;  // X3, X2, BA
    const float* gin;   // X2: gx2 | X1: gx1 | BA: gbA
    const float* gx2;   // BA only (pointwise)
    const float* x;     // X3 only (skip sums)
    const float* bB;    // X3
    const float* r1;    // X3, X2
    const float *wT0, *wL0, *wT1, *wL1;
    float *gwT0, *gwL0, *gwT1, *gwL1;
    int gw_assign;  // 1: this stage is the first writer of the gw buffers
    float* gz_out;
};

template <int TH, int TW>
struct BwdSmem {
    static constexpr int r4(int n) { return (n + 3) & ~3; }
    static constexpr int F12 = r4((TH + 12) * (TW + 12)), F4 = r4((TH + 4) * (TW + 4)), F2 = r4((TH + 2) * (TW + 2)),
                         F0 = r4(TH * TW);
    static constexpr int C6 = r4((TH / 2 + 6) * (TW / 2 + 6)), C4 = r4((TH / 2 + 4) * (TW / 2 + 4)),
                         C2 = r4((TH / 2 + 2) * (TW / 2 + 2)), C0 = r4((TH / 2) * (TW / 2));
    // zf,gA,gB | sA,sB,gl,goA,goB | lA,oB,oT,gsL,gsT | pz,gcA,gcB | sA1,sB1,gl1,goA1,goB1 | lA1,oB1,oT1,gsL1,gsT1 |
    // gzc | weights | gw accumulators | reduction scratch
    static constexpr int value = 3 * F12 + 5 * F4 + 5 * F2 + 3 * C6 + 5 * C4 + 5 * C2 + C0 + 8 * F4 + 8 * C4 +
                                 8 * F0 + 8 * C0 + 32 * 16 + 16;
};

// accumulate the 5 tap products of one (upstream, operand) pair into the four stats-parameter sums
__device__ __forceinline__ void stats_acc(float (&acc)[4], float ac, float ar, float ad, float au, float al) {
    acc[0] += ac;
    acc[1] += ar - ac;
    acc[2] += ad - ac;
    acc[3] += 4.f * ac - ar - ad - au - al;
}
// g[q] * y[q - o_t] for the five taps (y zero-extended)  -> St parameter gradient
__device__ __forceinline__ void stats_acc_St(float (&acc)[4], float g, const View& y, int h, int w) {
    const float* c = &y.at(h, w);
    stats_acc(acc, g * c[0], g * c[-1], g * c[-y.nw], g * c[y.nw], g * c[1]);
}
// gs[p] * z[p + o_t] (z clamp-extended) -> S parameter gradient
__device__ __forceinline__ void stats_acc_S(float (&acc)[4], float gs, const View& z, int h, int w) {
    const float* c = &z.at(h, w);
    stats_acc(acc, gs * c[0], gs * c[1], gs * c[z.nw], gs * c[-z.nw], gs * c[-1]);
}

// edge-weight gradient contributions of one pixel's four outgoing edges.
//   L     : gw_e -= gl * s[n_e]
//   GTV   : gw_e += D phi(t) + D w phi'(t) d,   D = go[p]-go[n], d = s[p]-s[n], t = w d  (linear: 2 w D d)
// also returns the d/dGamma sum for the thresholded case.
template <bool THR>
__device__ __forceinline__ float gtv_edge_grads(float* acc, int stride, const View& go, const View& s, const WViews& w,
                                                float G, int h, int x) {
    const float* cg = &go.at(h, x);
    const float* cs = &s.at(h, x);
    const int og[4] = {-go.nw, -1, 1, go.nw}, os[4] = {-s.nw, -1, 1, s.nw};
    float dG = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float we = w.e[e].at(h, x);
        float D = cg[0] - cg[og[e]], d = cs[0] - cs[os[e]];
        if (THR) {
            float t = we * d;
            acc[e * stride] += D * glr_phi(t, G) + D * we * glr_dphi(t, G) * d;
            if (fabsf(t) > G) dG += D * we * (t > 0.f ? -2.f : 2.f);
        } else {
            acc[e * stride] += 2.f * we * D * d;
        }
    }
    return dG;
}

template <int MODE, int TH, int TW>
__global__ void __launch_bounds__(512) k_block_bwd_stage(BlockBwdArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool HAS_A = MODE != BWD_BA;                  // adjoint of A(.) (GLR + linear GTV) on gA
    constexpr bool HAS_R = MODE == BWD_X2 || MODE == BWD_BA;  // adjoint of R(.) on gB
    constexpr bool THR = MODE == BWD_X2;                    // R is the thresholded one
    const int H = a.s.H, W = a.s.W, Hc = H / 2, Wc = W / 2, F = a.s.F, G = a.s.G;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW, hc0 = h0 / 2, wc0 = w0 / 2;
    const size_t HW = (size_t)H * W, HWc = (size_t)Hc * Wc;

    // ---- shared memory
    float* cur = smem;
    View zf = make_view(cur, h0 - 6, w0 - 6, TH + 12, TW + 12);
    View gA = make_view(cur, h0 - 6, w0 - 6, TH + 12, TW + 12);
    View gB = make_view(cur, h0 - 6, w0 - 6, TH + 12, TW + 12);
    View sA = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View sB = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View gl = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View goA = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View goB = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View lA = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View oB = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View oT = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View gsL = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View gsT = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View pz = make_view(cur, hc0 - 3, wc0 - 3, TH / 2 + 6, TW / 2 + 6);
    View gcA = make_view(cur, hc0 - 3, wc0 - 3, TH / 2 + 6, TW / 2 + 6);
    View gcB = make_view(cur, hc0 - 3, wc0 - 3, TH / 2 + 6, TW / 2 + 6);
    View sA1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View sB1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View gl1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View goA1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View goB1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View lA1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View oB1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View oT1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View gsL1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View gsT1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View gzc = make_view(cur, hc0, wc0, TH / 2, TW / 2);
    // the L adjoint reads the neighbours' weights, so wL needs the same (+)2 rectangle as wT here
    WViews wL0 = make_wviews(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    WViews wT0 = make_wviews(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    WViews wL1 = make_wviews(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    WViews wT1 = make_wviews(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    constexpr int NP = BwdSmem<TH, TW>::F0, NPc = BwdSmem<TH, TW>::C0;
    float* accL0 = cur; cur += 4 * NP;   // edge-weight gradient accumulators, [e][pixel of the tile]
    float* accT0 = cur; cur += 4 * NP;
    float* accL1 = cur; cur += 4 * NPc;
    float* accT1 = cur; cur += 4 * NPc;
    float* red = cur;

    // ---- per-graph scalars
    const float aT0 = expf(a.p.ro0[g]), aT1 = expf(a.p.ro1[g]);
    const float aL0 = expf(a.p.mu0[g]), aL1 = expf(a.p.mu1[g]);
    const float G0 = expf(a.p.gamma0[g]), G1 = expf(a.p.gamma1[g]);
    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const float s0 = a.p.skip ? a.p.skip[0] : 0.f, s1 = a.p.skip ? a.p.skip[1] : 1.f;
    const float c23 = al2 * s1;  // gr2 = c23 * gout

    // ---- weights, accumulators
    const size_t wplane = (size_t)plane * 4;
    tile_load_weights(wT0, a.wT0 + wplane * HW, H, W);
    tile_load_weights(wT1, a.wT1 + wplane * HWc, Hc, Wc);
    if (HAS_A) {
        tile_load_weights(wL0, a.wL0 + wplane * HW, H, W);
        tile_load_weights(wL1, a.wL1 + wplane * HWc, Hc, Wc);
    }
    TILE_LOOP(i, 8 * NP + 8 * NPc) accL0[i] = 0.f;

    // per-graph sums: 0 mu0, 1 ro0, 2 mu1, 3 ro1, 4 gamma0, 5 gamma1, 6 alpha_k, 7 beta2, 8 skip0, 9 skip1
    float gsum[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) gsum[k] = 0.f;

    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const size_t off = ((size_t)b * G * F + c) * HW;
        const StatsTaps kT0 = glr_load_taps(a.p.gtv0.stats, c), kT1 = glr_load_taps(a.p.gtv1.stats, c);
        const StatsTaps kL0 = glr_load_taps(a.p.glr0.stats, c), kL1 = glr_load_taps(a.p.glr1.stats, c);
        // per-channel stats sums: [module T0,L0,T1,L1][p01,p02a,p02b,p03]
        float st[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) st[k] = 0.f;
        float(&stT0)[4] = *reinterpret_cast<float(*)[4]>(st + 0);
        float(&stL0)[4] = *reinterpret_cast<float(*)[4]>(st + 4);
        float(&stT1)[4] = *reinterpret_cast<float(*)[4]>(st + 8);
        float(&stL1)[4] = *reinterpret_cast<float(*)[4]>(st + 12);

        __syncthreads();
        // ---- phase 1: stage input (clamp-extended) and upstream gradients (zero-extended)
        tile_load_clamped(zf, a.z + off, H, W);
        TILE_LOOP(i, gA.size()) {
            int h = gA.oh + i / gA.nw, w = gA.ow + i % gA.nw;
            float va = 0.f, vb = 0.f;
            if (glr_inside(h, w, H, W)) {
                const size_t gi = off + (size_t)h * W + w;
                if (MODE == BWD_X3) {
                    va = -c23 * a.gout[gi];
                } else if (MODE == BWD_X2) {
                    float gr2 = c23 * a.gout[gi];
                    float gr1 = be2 * gr2 + al1 * a.gin[gi];
                    va = -gr1;
                    vb = gr2 + gr1;
                } else if (MODE == BWD_X1) {
                    va = -al0 * a.gin[gi];
                } else {
                    vb = a.gin[gi];
                }
            }
            if (HAS_A) gA.p[i] = va;
            if (HAS_R) gB.p[i] = vb;
        }
        __syncthreads();
        // ---- phase 2: forward S + pooling; adjoint of St on the upstreams; pooled upstreams
        if (HAS_A) tile_S2(sA, kL0, sB, kT0, zf, H, W);
        else tile_S(sB, zf, kT0, H, W);
        tile_pool(pz, zf, Hc, Wc);
        if (HAS_A) {
            tile_Szero<false>(gl, gA, kL0, aL0, H, W);
            tile_Szero<true>(goA, gA, kT0, aT0, H, W);
        }
        if (HAS_R) tile_Szero<true>(goB, gB, kT0, aT0, H, W);
        TILE_LOOP(i, gcA.size()) {  // VJP of P^T is P: mean of the 2x2 block, zero outside the coarse image
            int h = gcA.oh + i / gcA.nw, w = gcA.ow + i % gcA.nw;
            float va = 0.f, vb = 0.f;
            if (glr_inside(h, w, Hc, Wc)) {
                if (HAS_A) { const float* q = &gA.at(2 * h, 2 * w); va = 0.25f * (q[0] + q[1] + q[gA.nw] + q[gA.nw + 1]); }
                if (HAS_R) { const float* q = &gB.at(2 * h, 2 * w); vb = 0.25f * (q[0] + q[1] + q[gB.nw] + q[gB.nw + 1]); }
            }
            if (HAS_A) gcA.p[i] = va;
            if (HAS_R) gcB.p[i] = vb;
        }
        __syncthreads();
        // ---- phase 3: forward cores (fine), coarse S; adjoint cores (fine), coarse St-adjoints
        if (HAS_A) tile_L(lA, sA, wL0, H, W);
        tile_gtv_core<false>(oB, sB, wT0, 0.f, H, W);
        if (THR) tile_gtv_core<true>(oT, sB, wT0, G0, H, W);
        if (HAS_A) tile_S2(sA1, kL1, sB1, kT1, pz, Hc, Wc);
        else tile_S(sB1, pz, kT1, Hc, Wc);
        if (HAS_A) {
            tile_L_adj(gsL, gl, wL0, H, W);
            tile_gtv_core<false>(gsT, goA, wT0, 0.f, H, W);  // the linear core is self-adjoint
            if (THR) tile_gtv_core_thr_adj<true>(gsT, goB, sB, wT0, G0, H, W);
            tile_Szero<false>(gl1, gcA, kL1, aL1, Hc, Wc);
            tile_Szero<true>(goA1, gcA, kT1, aT1, Hc, Wc);
        } else {
            tile_gtv_core<false>(gsT, goB, wT0, 0.f, H, W);
        }
        if (HAS_R) tile_Szero<true>(goB1, gcB, kT1, aT1, Hc, Wc);
        __syncthreads();
        // ---- phase 4: coarse cores, forward and adjoint
        if (HAS_A) tile_L(lA1, sA1, wL1, Hc, Wc);
        tile_gtv_core<false>(oB1, sB1, wT1, 0.f, Hc, Wc);
        if (THR) tile_gtv_core<true>(oT1, sB1, wT1, G1, Hc, Wc);
        if (HAS_A) {
            tile_L_adj(gsL1, gl1, wL1, Hc, Wc);
            tile_gtv_core<false>(gsT1, goA1, wT1, 0.f, Hc, Wc);
            if (THR) tile_gtv_core_thr_adj<true>(gsT1, goB1, sB1, wT1, G1, Hc, Wc);
        } else {
            tile_gtv_core<false>(gsT1, goB1, wT1, 0.f, Hc, Wc);
        }
        __syncthreads();
        // ---- phase 5: coarse epilogue: gradient wrt P z, parameter sums, coarse edge-weight gradients
        TILE_LOOP(i, gzc.size()) {
            const int lh = i / gzc.nw, lw = i % gzc.nw, h = hc0 + lh, w = wc0 + lw;
            if (h >= Hc || w >= Wc) { gzc.p[i] = 0.f; continue; }
            float v = tile_S_adj_at(gsT1, kT1, h, w, Hc, Wc);
            if (HAS_A) v += tile_S_adj_at(gsL1, kL1, h, w, Hc, Wc);
            gzc.p[i] = v;
            const float ga = HAS_A ? gcA.at(h, w) : 0.f, gb = HAS_R ? gcB.at(h, w) : 0.f;
            // forward values for the mu1 / ro1 gradients
            const float gtv_lin = tile_St_at(oB1, kT1, h, w);
            if (HAS_A) {
                gsum[2] += aL1 * ga * tile_St_at(lA1, kL1, h, w);
                gsum[3] += aT1 * ga * gtv_lin;
                stats_acc_St(stL1, aL1 * ga, lA1, h, w);
                stats_acc_S(stL1, gsL1.at(h, w), pz, h, w);
                stats_acc_St(stT1, aT1 * ga, oB1, h, w);
            }
            if (HAS_R) {
                gsum[3] += aT1 * gb * (THR ? tile_St_at(oT1, kT1, h, w) : gtv_lin);
                stats_acc_St(stT1, aT1 * gb, THR ? oT1 : oB1, h, w);
            }
            stats_acc_S(stT1, gsT1.at(h, w), pz, h, w);
            // edge weights
            const int pi = lh * (TW / 2) + lw;
            if (HAS_A) {
                const float glv = gl1.at(h, w);
                const float* cs = &sA1.at(h, w);
                accL1[0 * NPc + pi] -= glv * cs[-sA1.nw];
                accL1[1 * NPc + pi] -= glv * cs[-1];
                accL1[2 * NPc + pi] -= glv * cs[1];
                accL1[3 * NPc + pi] -= glv * cs[sA1.nw];
                gtv_edge_grads<false>(accT1 + pi, NPc, goA1, sB1, wT1, 0.f, h, w);
            }
            if (HAS_R) {
                if (THR) gsum[5] += gtv_edge_grads<true>(accT1 + pi, NPc, goB1, sB1, wT1, G1, h, w);
                else gtv_edge_grads<false>(accT1 + pi, NPc, goB1, sB1, wT1, 0.f, h, w);
            }
        }
        __syncthreads();
        // ---- phase 6: fine epilogue
        TILE_LOOP(i, TH * TW) {
            const int lh = i / TW, lw = i % TW, h = h0 + lh, w = w0 + lw;
            if (h >= H || w >= W) continue;
            const size_t gi = off + (size_t)h * W + w;
            const float zv = zf.at(h, w);
            const float ga = HAS_A ? gA.at(h, w) : 0.f, gb = HAS_R ? gB.at(h, w) : 0.f;
            float V = ga + tile_S_adj_at(gsT, kT0, h, w, H, W) + 0.25f * gzc.at(h >> 1, w >> 1);
            if (HAS_A) V += tile_S_adj_at(gsL, kL0, h, w, H, W);
            // forward values
            const float gtv_lin = tile_St_at(oB, kT0, h, w);
            float glr = 0.f;
            if (HAS_A) {
                glr = tile_St_at(lA, kL0, h, w);
                gsum[0] += aL0 * ga * glr;
                gsum[1] += aT0 * ga * gtv_lin;
                stats_acc_St(stL0, aL0 * ga, lA, h, w);
                stats_acc_S(stL0, gsL.at(h, w), zf, h, w);
                stats_acc_St(stT0, aT0 * ga, oB, h, w);
            }
            if (HAS_R) {
                gsum[1] += aT0 * gb * (THR ? tile_St_at(oT, kT0, h, w) : gtv_lin);
                stats_acc_St(stT0, aT0 * gb, THR ? oT : oB, h, w);
            }
            stats_acc_S(stT0, gsT.at(h, w), zf, h, w);
            // edge weights
            const int pi = lh * TW + lw;
            if (HAS_A) {
                const float glv = gl.at(h, w);
                const float* cs = &sA.at(h, w);
                accL0[0 * NP + pi] -= glv * cs[-sA.nw];
                accL0[1 * NP + pi] -= glv * cs[-1];
                accL0[2 * NP + pi] -= glv * cs[1];
                accL0[3 * NP + pi] -= glv * cs[sA.nw];
                gtv_edge_grads<false>(accT0 + pi, NP, goA, sB, wT0, 0.f, h, w);
            }
            if (HAS_R) {
                if (THR) gsum[4] += gtv_edge_grads<true>(accT0 + pi, NP, goB, sB, wT0, G0, h, w);
                else gtv_edge_grads<false>(accT0 + pi, NP, goB, sB, wT0, 0.f, h, w);
            }
            // stage epilogue
            (void)zv;
            float outv;
            if (MODE == BWD_X3) {
                const float go_ = a.gout[gi];
                outv = s1 * go_ + V;
            } else if (MODE == BWD_X2) {
                const float gx2v = a.gin[gi];
                outv = gx2v + V;
                gsum[6] += gx2v * a.r1[gi];
            } else if (MODE == BWD_X1) {
                outv = (1.f + al0) * a.gin[gi] + V;
            } else {
                const float go_ = a.gout[gi];
                const float gr2 = c23 * go_;
                const float gbB = gr2 + be2 * gr2 + al1 * a.gx2[gi];
                outv = a.gin[gi] + V + gbB + s0 * go_;
            }
            a.gz_out[gi] = outv;
        }
        // X3 / X1 also need A(z) itself (r2, u2, x3 / bA - A bA): a second pass once the coarse forward term is known
        if (MODE == BWD_X3 || MODE == BWD_X1) {
            TILE_LOOP(i, TH * TW) {
                const int h = h0 + i / TW, w = w0 + i % TW;
                if (h >= H || w >= W) continue;
                const size_t gi = off + (size_t)h * W + w;
                const float zv = zf.at(h, w);
                const int hc = h >> 1, wc = w >> 1;
                const float tcv = aL1 * tile_St_at(lA1, kL1, hc, wc) + aT1 * tile_St_at(oB1, kT1, hc, wc);
                const float Az = zv + aL0 * tile_St_at(lA, kL0, h, w) + aT0 * tile_St_at(oB, kT0, h, w) + 0.25f * tcv;
                if (MODE == BWD_X3) {
                    const float r1v = a.r1[gi], go_ = a.gout[gi];
                    const float u2 = (a.bB[gi] - Az) + be2 * r1v;
                    const float x3 = zv + al2 * u2;
                    const float g3 = s1 * go_;
                    gsum[6] += g3 * u2;
                    gsum[7] += al2 * g3 * r1v;
                    if (a.p.skip) { gsum[8] += go_ * a.x[gi]; gsum[9] += go_ * x3; }
                } else {
                    gsum[6] += a.gin[gi] * (zv - Az);
                }
            }
        }
        // ---- per-channel stats gradients
        block_sum_n<16>(st, red);
        if (threadIdx.x == 0) {
            const int C = G * F;
            float* dst[4] = {a.gr.gtv0_stats, a.gr.glr0_stats, a.gr.gtv1_stats, a.gr.glr1_stats};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (!HAS_A && (m == 1 || m == 3)) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k) atomicAdd(&dst[m][k * C + c], st[m * 4 + k]);
            }
        }
    }

    // ---- edge-weight gradients of this tile: one read-modify-write per stage
    __syncthreads();
    {
        float* gw0[2] = {a.gwT0 + wplane * HW, a.gwL0 + wplane * HW};
        const float* acc0[2] = {accT0, accL0};
        for (int m = 0; m < (HAS_A ? 2 : 1); ++m)
            TILE_LOOP(i, 4 * TH * TW) {
                const int e = i / (TH * TW), pi = i % (TH * TW), h = h0 + pi / TW, w = w0 + pi % TW;
                if (h >= H || w >= W) continue;
                float* q = gw0[m] + (size_t)e * HW + (size_t)h * W + w;
                const float v = acc0[m][e * NP + pi];
                *q = a.gw_assign ? v : *q + v;
            }
        float* gw1[2] = {a.gwT1 + wplane * HWc, a.gwL1 + wplane * HWc};
        const float* acc1[2] = {accT1, accL1};
        for (int m = 0; m < (HAS_A ? 2 : 1); ++m)
            TILE_LOOP(i, TH * TW) {  // 4 * (TH/2) * (TW/2)
                const int e = i / (TH * TW / 4), pi = i % (TH * TW / 4), h = hc0 + pi / (TW / 2), w = wc0 + pi % (TW / 2);
                if (h >= Hc || w >= Wc) continue;
                float* q = gw1[m] + (size_t)e * HWc + (size_t)h * Wc + w;
                const float v = acc1[m][e * NPc + pi];
                *q = a.gw_assign ? v : *q + v;
            }
    }
    // ---- per-graph scalar gradients
    block_sum_n<10>(gsum, red);
    if (threadIdx.x == 0) {
        if (HAS_A) {
            atomicAdd(&a.gr.mu0[g], gsum[0]);
            atomicAdd(&a.gr.mu1[g], gsum[2]);
        }
        atomicAdd(&a.gr.ro0[g], gsum[1]);
        atomicAdd(&a.gr.ro1[g], gsum[3]);
        if (THR) {
            atomicAdd(&a.gr.gamma0[g], gsum[4] * G0);
            atomicAdd(&a.gr.gamma1[g], gsum[5] * G1);
        }
        if (MODE == BWD_X3) {
            atomicAdd(&a.gr.alpha[2 * G + g], gsum[6]);
            atomicAdd(&a.gr.beta[2 * G + g], gsum[7]);
            if (a.p.skip && a.gr.skip) {
                atomicAdd(&a.gr.skip[0], gsum[8]);
                atomicAdd(&a.gr.skip[1], gsum[9]);
            }
        }
        if (MODE == BWD_X2) atomicAdd(&a.gr.alpha[G + g], gsum[6]);
        if (MODE == BWD_X1) atomicAdd(&a.gr.alpha[g], gsum[6]);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#ifndef GLR_BTH
#define GLR_BTH 32
#define GLR_BTW 32
#endif
#define GLR_BWD_THREADS 512

template <int MODE>
static int launch_bwd_stage(const BlockBwdArgs& a, void* stream) {
    const glrgtv_shape& s = a.s;
    const long tiles = (long)((s.W + GLR_BTW - 1) / GLR_BTW) * ((s.H + GLR_BTH - 1) / GLR_BTH);
    const long blocks = tiles * s.B * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    constexpr size_t smem = BwdSmem<GLR_BTH, GLR_BTW>::value * sizeof(float);
    static_assert(smem <= 227 * 1024, "backward tile does not fit shared memory");
#ifndef GLRGTV_EMU
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_block_bwd_stage<MODE, GLR_BTH, GLR_BTW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return glr_record_launch_error();
        configured = true;
    }
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_X3 + MODE, stream);
    GLR_LAUNCH((k_block_bwd_stage<MODE, GLR_BTH, GLR_BTW>), dim3((unsigned)blocks), GLR_BWD_THREADS, smem, stream, a);
    GLR_PROF_END(GLRGTV_SLOT_BWD_X3 + MODE, stream);
    return GLR_CHECK_LAUNCH();
}

// strided edge-weight backward (ops_basic.cu)
int glr_edge_weights_bwd_strided(const glrgtv_shape* s, const glrgtv_window* win, const float* feat, size_t feat_bs,
                                 const float* multiM, const float* w, const float* gw, float* gfeat, size_t gfeat_bs,
                                 float* gmultiM, float* scratch, void* stream);

static size_t ws_floats(const glrgtv_shape* s, size_t* o_gx2, size_t* o_gx1, size_t* o_gbA, size_t* o_gw, size_t* o_scr) {
    const size_t N = (size_t)s->B * s->H * s->W, C = (size_t)s->G * s->F, GE = (size_t)s->G * 4;
    size_t off = 0;
    *o_gx2 = off; off += C * N;
    *o_gx1 = off; off += C * N;
    *o_gbA = off; off += C * N;
    *o_gw = off;  off += 2 * GE * N + 2 * GE * (N / 4);
    *o_scr = off; off += (size_t)s->G * 5 * N;
    return off;
}

extern "C" size_t glrgtv_block_bwd_workspace_bytes(const glrgtv_shape* s) {
    if (!glr_shape_ok(s)) return 0;
    size_t a, b, c, d, e;
    return ws_floats(s, &a, &b, &c, &d, &e) * sizeof(float);
}

int glr_block_params_ok(const glrgtv_shape* s, const glrgtv_block_params* p);

extern "C" int glrgtv_block_bwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                                const float* feat0, const float* feat1, const glrgtv_block_saved* sv,
                                const float* gout, float* gx, float* gfeat0, float* gfeat1,
                                const glrgtv_block_grads* gr, void* workspace, size_t workspace_bytes, void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    int rc = glr_block_params_ok(s, p);
    if (rc) return rc;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(feat0); GLR_REQUIRE_PTR(feat1); GLR_REQUIRE_PTR(gout);
    GLR_REQUIRE_PTR(gx); GLR_REQUIRE_PTR(gfeat0); GLR_REQUIRE_PTR(gfeat1); GLR_REQUIRE_PTR(workspace);
    if (!sv || !gr) return GLRGTV_ERR_POINTER;
    const float* need[9] = {sv->wT0, sv->wL0, sv->wT1, sv->wL1, sv->bA, sv->x1, sv->bB, sv->r1, sv->x2};
    for (int i = 0; i < 9; ++i) GLR_REQUIRE_PTR(need[i]);
    float* const* gp = &gr->gtv0_stats;
    for (int i = 0; i < 16; ++i) GLR_REQUIRE_PTR(gp[i]);
    if (p->skip) GLR_REQUIRE_PTR(gr->skip);
    size_t o_gx2, o_gx1, o_gbA, o_gw, o_scr;
    if (workspace_bytes < ws_floats(s, &o_gx2, &o_gx1, &o_gbA, &o_gw, &o_scr) * sizeof(float)) return GLRGTV_ERR_WORKSPACE;

    float* ws = (float*)workspace;
    const size_t N = (size_t)s->B * s->H * s->W, GE = (size_t)s->G * 4;
    BlockBwdArgs a;
    a.s = *s; a.p = *p; a.gr = *gr;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1;
    a.gwT0 = ws + o_gw; a.gwL0 = a.gwT0 + GE * N; a.gwT1 = a.gwL0 + GE * N; a.gwL1 = a.gwT1 + GE * (N / 4);
    a.gout = gout; a.x = x; a.bB = sv->bB; a.r1 = sv->r1; a.gx2 = ws + o_gx2;

    a.z = sv->x2; a.gin = nullptr; a.gz_out = ws + o_gx2; a.gw_assign = 1;
    if ((rc = launch_bwd_stage<BWD_X3>(a, stream))) return rc;
    a.z = sv->x1; a.gin = ws + o_gx2; a.gz_out = ws + o_gx1; a.gw_assign = 0;
    if ((rc = launch_bwd_stage<BWD_X2>(a, stream))) return rc;
    a.z = sv->bA; a.gin = ws + o_gx1; a.gz_out = ws + o_gbA;
    if ((rc = launch_bwd_stage<BWD_X1>(a, stream))) return rc;
    a.z = x; a.gin = ws + o_gbA; a.gz_out = gx;
    if ((rc = launch_bwd_stage<BWD_BA>(a, stream))) return rc;

    // edge weights -> features (four sets; feat halves are strided in the batch dimension)
    glrgtv_window win;
    win.n_edges = 4;
    const int dh[4] = {-1, 0, 0, 1}, dw[4] = {0, -1, 1, 0};
    for (int e = 0; e < 4; ++e) { win.dh[e] = dh[e]; win.dw[e] = dw[e]; }
    glrgtv_shape sc = *s;
    sc.H /= 2; sc.W /= 2;
    const size_t C = (size_t)s->G * s->F, HW = (size_t)s->H * s->W, HWc = HW / 4;
    float* scr = ws + o_scr;
    GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_WEIGHTS, stream);
    if ((rc = glr_edge_weights_bwd_strided(s, &win, feat0, 2 * C * HW, p->gtv0.multiM, sv->wT0, a.gwT0, gfeat0,
                                           2 * C * HW, gr->gtv0_M, scr, stream))) return rc;
    if ((rc = glr_edge_weights_bwd_strided(s, &win, feat0 + C * HW, 2 * C * HW, p->glr0.multiM, sv->wL0, a.gwL0,
                                           gfeat0 + C * HW, 2 * C * HW, gr->glr0_M, scr, stream))) return rc;
    if ((rc = glr_edge_weights_bwd_strided(&sc, &win, feat1, 2 * C * HWc, p->gtv1.multiM, sv->wT1, a.gwT1, gfeat1,
                                           2 * C * HWc, gr->gtv1_M, scr, stream))) return rc;
    rc = glr_edge_weights_bwd_strided(&sc, &win, feat1 + C * HWc, 2 * C * HWc, p->glr1.multiM, sv->wL1, a.gwL1,
                                      gfeat1 + C * HWc, 2 * C * HWc, gr->glr1_M, scr, stream);
    GLR_PROF_END(GLRGTV_SLOT_BWD_WEIGHTS, stream);
    return rc;
}
