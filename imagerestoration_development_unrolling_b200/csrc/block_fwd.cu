// block_fwd.cu - fused forward of LocalLowpassFilteringBlock / MixtureGTVGLR (V1X0:707-811, 985-988).
//
// Staged fusion: the unrolled solver is cut where a global dependency forces it, and each
// stage is ONE kernel that reads its operands once from HBM, keeps the whole operator chain of
// A(.) = I + e^mu0 St L S + e^ro0 St Ct C S + P^T[ e^mu1 ... + e^ro1 ... ]P  (V1X0:642-682)
// in shared memory, and writes only what later stages or the backward pass need:
//
//   k_block_weights x2   feat0 / feat1           -> wT0,wL0 / wT1,wL1        (V1X0:716-733)
//   MODE_BA              y                        -> bA = y + R_lin(y)         (V1X0:738-749)
//   MODE_X1              bA                       -> x1 = bA + a0 (bA - A bA)  (V1X0:751-753)
//   MODE_X2              x1, y                    -> bB, r1, x2                (V1X0:757-786)
//   MODE_X3              x2, bB, r1, x            -> out                       (V1X0:788-790, 985-988)
//
// with R(z) = e^ro0 St0 Ct0 phi(C0 S0 z) + P^T e^ro1 St1 Ct1 phi(C1 S1 P z), phi = identity (R_lin) or
// 2*soft-threshold - id (bB).  Tile = TH x TW fine pixels, input halo 6 (3 for the fine chain, 3 coarse
// pixels = 6 fine for the half-resolution chain).
#include "tile.cuh"

enum { MODE_BA = 0, MODE_X1 = 1, MODE_X2 = 2, MODE_X3 = 3 };

struct BlockFwdArgs {
    glrgtv_shape s;
    glrgtv_block_params p;
    const float* z;      // stencil input: y | bA | x1 | x2
    const float* y;      // X2: y ; X3: x (skip path)
    const float* bB_in;  // X3
    const float* r1_in;  // X3
    const float *wT0, *wL0, *wT1, *wL1;
    float* out0;  // BA: bA | X1: x1 | X2: x2 | X3: out
    float* out1;  // X2: bB
    float* out2;  // X2: r1
};

template <int TH, int TW>
struct FwdSmem {
    // number of floats the kernel carves (upper bound over modes), each view rounded up to 4 floats
    static constexpr int r4(int n) { return (n + 3) & ~3; }
    static constexpr int value =
        r4((TH + 12) * (TW + 12)) + 2 * r4((TH + 4) * (TW + 4)) + 3 * r4((TH + 2) * (TW + 2)) +
        r4((TH / 2 + 6) * (TW / 2 + 6)) + 2 * r4((TH / 2 + 4) * (TW / 2 + 4)) + 3 * r4((TH / 2 + 2) * (TW / 2 + 2)) +
        2 * r4((TH / 2) * (TW / 2)) + 4 * r4((TH + 2) * (TW + 2)) + 4 * r4((TH + 4) * (TW + 4)) +
        4 * r4((TH / 2 + 2) * (TW / 2 + 2)) + 4 * r4((TH / 2 + 4) * (TW / 2 + 4)) + 32;
};

template <int MODE, int TH, int TW>
__global__ void __launch_bounds__(GLR_THREADS) k_block_stage(BlockFwdArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool GLR = MODE != MODE_BA;  // BA only needs the GTV chain
    constexpr bool THR = MODE == MODE_X2;  // thresholded right-hand side
    const int H = a.s.H, W = a.s.W, Hc = H / 2, Wc = W / 2, F = a.s.F, G = a.s.G;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW;
    const int hc0 = h0 / 2, wc0 = w0 / 2;
    const size_t HW = (size_t)H * W, HWc = (size_t)Hc * Wc;

    // ---- carve shared memory
    float* cur = smem;
    View zf = make_view(cur, h0 - 6, w0 - 6, TH + 12, TW + 12);
    View sA = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View sB = make_view(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    View lA = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View oB = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View oT = make_view(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    View pz = make_view(cur, hc0 - 3, wc0 - 3, TH / 2 + 6, TW / 2 + 6);
    View sA1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View sB1 = make_view(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);
    View lA1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View oB1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View oT1 = make_view(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    View tc = make_view(cur, hc0, wc0, TH / 2, TW / 2);
    View tcT = make_view(cur, hc0, wc0, TH / 2, TW / 2);
    WViews wL0 = make_wviews(cur, h0 - 1, w0 - 1, TH + 2, TW + 2);
    WViews wT0 = make_wviews(cur, h0 - 2, w0 - 2, TH + 4, TW + 4);
    WViews wL1 = make_wviews(cur, hc0 - 1, wc0 - 1, TH / 2 + 2, TW / 2 + 2);
    WViews wT1 = make_wviews(cur, hc0 - 2, wc0 - 2, TH / 2 + 4, TW / 2 + 4);

    // ---- per-graph scalars
    const float aT0 = expf(a.p.ro0[g]), aT1 = expf(a.p.ro1[g]);
    const float aL0 = GLR ? expf(a.p.mu0[g]) : 0.f, aL1 = GLR ? expf(a.p.mu1[g]) : 0.f;
    const float G0 = THR ? expf(a.p.gamma0[g]) : 0.f, G1 = THR ? expf(a.p.gamma1[g]) : 0.f;
    float alpha = 0.f, beta2 = 0.f, s0 = 0.f, s1 = 1.f;
    if (MODE == MODE_X1) alpha = a.p.alpha[0 * G + g];
    if (MODE == MODE_X2) alpha = a.p.alpha[1 * G + g];
    if (MODE == MODE_X3) {
        alpha = a.p.alpha[2 * G + g];
        beta2 = a.p.beta[2 * G + g];
        if (a.p.skip) { s0 = a.p.skip[0]; s1 = a.p.skip[1]; }
    }

    // ---- weights of this graph (shared by its F channels)
    const size_t wplane = (size_t)plane * 4;
    tile_load_weights(wT0, a.wT0 + wplane * HW, H, W);
    tile_load_weights(wT1, a.wT1 + wplane * HWc, Hc, Wc);
    if (GLR) {
        tile_load_weights(wL0, a.wL0 + wplane * HW, H, W);
        tile_load_weights(wL1, a.wL1 + wplane * HWc, Hc, Wc);
    }

    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const size_t off = ((size_t)b * G * F + c) * HW;
        const StatsTaps kT0 = glr_load_taps(a.p.gtv0.stats, c), kT1 = glr_load_taps(a.p.gtv1.stats, c);
        StatsTaps kL0 = kT0, kL1 = kT1;
        if (GLR) { kL0 = glr_load_taps(a.p.glr0.stats, c); kL1 = glr_load_taps(a.p.glr1.stats, c); }

        __syncthreads();  // previous channel's epilogue is done with the views
        tile_load_clamped(zf, a.z + off, H, W);
        __syncthreads();
        // phase 1: S on the fine grid, pooling
        if (GLR) tile_S2(sA, kL0, sB, kT0, zf, H, W);
        else tile_S(sB, zf, kT0, H, W);
        tile_pool(pz, zf, Hc, Wc);
        __syncthreads();
        // phase 2: fine L / GTV cores, coarse S
        if (GLR) tile_L(lA, sA, wL0, H, W);
        tile_gtv_core<false>(oB, sB, wT0, 0.f, H, W);
        if (THR) tile_gtv_core<true>(oT, sB, wT0, G0, H, W);
        if (GLR) tile_S2(sA1, kL1, sB1, kT1, pz, Hc, Wc);
        else tile_S(sB1, pz, kT1, Hc, Wc);
        __syncthreads();
        // phase 3: coarse cores
        if (GLR) tile_L(lA1, sA1, wL1, Hc, Wc);
        tile_gtv_core<false>(oB1, sB1, wT1, 0.f, Hc, Wc);
        if (THR) tile_gtv_core<true>(oT1, sB1, wT1, G1, Hc, Wc);
        __syncthreads();
        // phase 4: coarse St, scaled
        TILE_LOOP(i, tc.size()) {
            int h = hc0 + i / tc.nw, w = wc0 + i % tc.nw;
            float v = aT1 * tile_St_at(oB1, kT1, h, w);
            if (GLR) v += aL1 * tile_St_at(lA1, kL1, h, w);
            tc.p[i] = v;
            if (THR) tcT.p[i] = aT1 * tile_St_at(oT1, kT1, h, w);
        }
        __syncthreads();
        // phase 5: fine St + epilogue
        TILE_LOOP(i, TH * TW) {
            int h = h0 + i / TW, w = w0 + i % TW;
            if (h >= H || w >= W) continue;
            const size_t gi = off + (size_t)h * W + w;
            const float zv = zf.at(h, w);
            float Az = zv + aT0 * tile_St_at(oB, kT0, h, w) + 0.25f * tc.at(h >> 1, w >> 1);
            if (GLR) Az += aL0 * tile_St_at(lA, kL0, h, w);
            if (MODE == MODE_BA) {
                a.out0[gi] = Az;  // y + R_lin(y)
            } else if (MODE == MODE_X1) {
                a.out0[gi] = zv + alpha * (zv - Az);
            } else if (MODE == MODE_X2) {
                float bB = a.y[gi] + aT0 * tile_St_at(oT, kT0, h, w) + 0.25f * tcT.at(h >> 1, w >> 1);
                float r1 = bB - Az;
                a.out1[gi] = bB;
                a.out2[gi] = r1;
                a.out0[gi] = zv + alpha * r1;
            } else {
                float r1 = a.r1_in[gi];
                float u2 = (a.bB_in[gi] - Az) + beta2 * r1;
                float x3 = zv + alpha * u2;
                a.out0[gi] = a.p.skip ? s0 * a.y[gi] + s1 * x3 : x3;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// edge weights of both operator families of one scale in one launch (V1X0:716-733).
// feat [B, 2C, H, W]: channels [0,C) -> GTV weights, [C,2C) -> GLR weights.  plane = (b, set, g).
// ---------------------------------------------------------------------------------------------------
template <int TH, int TW>
__global__ void __launch_bounds__(GLR_THREADS) k_block_weights(glrgtv_shape s, const float* __restrict__ feat,
                                                              const float* __restrict__ M_gtv,
                                                              const float* __restrict__ M_glr,
                                                              float* __restrict__ w_gtv, float* __restrict__ w_glr) {
    GLR_SMEM_DECL(smem);
    const int H = s.H, W = s.W, F = s.F, G = s.G, C = G * F;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, set = (plane / G) % 2, b = plane / (2 * G);
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW;
    const size_t HW = (size_t)H * W;
    constexpr int NH = TH + 2, NW = TW + 2, NP = NH * NW;
    const float* fp = feat + ((size_t)b * 2 * C + (size_t)set * C + (size_t)g * F) * HW;
    const float* Mg = (set ? M_glr : M_gtv) + g * F;
    float* wp = (set ? w_glr : w_gtv) + ((size_t)b * G + g) * 4 * HW;
    // phase 1: normalised, scaled features on the tile (+) 1 (clamp-extended), layout ft[f][pixel]
    TILE_LOOP(i, NP) {
        int h = glr_clampi(h0 - 1 + i / NW, 0, H - 1), w = glr_clampi(w0 - 1 + i % NW, 0, W - 1);
        const float* q = fp + (size_t)h * W + w;
        float n2 = 0.f;
        for (int f = 0; f < F; ++f) { float v = q[f * HW]; n2 += v * v; }
        float nrm = fmaxf(sqrtf(n2), 1e-12f);
        for (int f = 0; f < F; ++f) smem[f * NP + i] = q[f * HW] / nrm * Mg[f];
    }
    __syncthreads();
    // phase 2: similarities with the four neighbours, softmax
    TILE_LOOP(i, TH * TW) {
        int lh = i / TW, lw = i % TW, h = h0 + lh, w = w0 + lw;
        if (h >= H || w >= W) continue;
        const float* c = smem + (lh + 1) * NW + (lw + 1);
        float su = 0.f, sl = 0.f, sr = 0.f, sd = 0.f;
        for (int f = 0; f < F; ++f) {
            const float* cf = c + f * NP;
            float v = cf[0];
            su += v * cf[-NW]; sl += v * cf[-1]; sr += v * cf[1]; sd += v * cf[NW];
        }
        float mx = fmaxf(fmaxf(su, sl), fmaxf(sr, sd));
        su = expf(su - mx); sl = expf(sl - mx); sr = expf(sr - mx); sd = expf(sd - mx);
        float inv = 1.f / (su + sl + sr + sd);
        size_t gi = (size_t)h * W + w;
        wp[gi] = su * inv; wp[HW + gi] = sl * inv; wp[2 * HW + gi] = sr * inv; wp[3 * HW + gi] = sd * inv;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#ifndef GLR_TH
#define GLR_TH 32
#define GLR_TW 32
#endif
#define GLR_WT_TH 16
#define GLR_WT_TW 32

template <int MODE>
static int launch_stage(const BlockFwdArgs& a, void* stream) {
    const glrgtv_shape& s = a.s;
    const long tiles = (long)((s.W + GLR_TW - 1) / GLR_TW) * ((s.H + GLR_TH - 1) / GLR_TH);
    const long blocks = tiles * s.B * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    constexpr size_t smem = FwdSmem<GLR_TH, GLR_TW>::value * sizeof(float);
#ifndef GLRGTV_EMU
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_block_stage<MODE, GLR_TH, GLR_TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return glr_record_launch_error();
        configured = true;
    }
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_BA + MODE, stream);
    GLR_LAUNCH((k_block_stage<MODE, GLR_TH, GLR_TW>), dim3((unsigned)blocks), GLR_THREADS, smem, stream, a);
    GLR_PROF_END(GLRGTV_SLOT_FWD_BA + MODE, stream);
    return GLR_CHECK_LAUNCH();
}

static int launch_weights(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, float* wt,
                          float* wl, void* stream) {
    const long tiles = (long)((s.W + GLR_WT_TW - 1) / GLR_WT_TW) * ((s.H + GLR_WT_TH - 1) / GLR_WT_TH);
    const long blocks = tiles * s.B * 2 * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    const size_t smem = (size_t)s.F * (GLR_WT_TH + 2) * (GLR_WT_TW + 2) * sizeof(float);
    if (smem > 200 * 1024) return GLRGTV_ERR_UNSUPPORTED;
#ifndef GLRGTV_EMU
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k_block_weights<GLR_WT_TH, GLR_WT_TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return glr_record_launch_error();
    }
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    GLR_LAUNCH((k_block_weights<GLR_WT_TH, GLR_WT_TW>), dim3((unsigned)blocks), GLR_THREADS, smem, stream, s, feat, Mt,
               Ml, wt, wl);
    GLR_PROF_END(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    return GLR_CHECK_LAUNCH();
}

int glr_block_params_ok(const glrgtv_shape* s, const glrgtv_block_params* p) {
    if (!p) return GLRGTV_ERR_POINTER;
    const glrgtv_opparams* ops[4] = {&p->gtv0, &p->glr0, &p->gtv1, &p->glr1};
    for (int i = 0; i < 4; ++i) {
        const glrgtv_stats& st = ops[i]->stats;
        if (!glr_aligned(st.p01) || !glr_aligned(st.p02a) || !glr_aligned(st.p02b) || !glr_aligned(st.p03) ||
            !glr_aligned(ops[i]->multiM))
            return GLRGTV_ERR_POINTER;
        if (st.n != s->G * s->F || st.pad != GLRGTV_PAD_CLAMP) return GLRGTV_ERR_SHAPE;
    }
    const float* v[8] = {p->alpha, p->beta, p->mu0, p->ro0, p->gamma0, p->mu1, p->ro1, p->gamma1};
    for (int i = 0; i < 8; ++i)
        if (!glr_aligned(v[i])) return GLRGTV_ERR_POINTER;
    if (p->skip && !glr_aligned(p->skip)) return GLRGTV_ERR_POINTER;
    return GLRGTV_OK;
}

extern "C" int glrgtv_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                                const float* feat0, const float* feat1, float* out, const glrgtv_block_saved* sv,
                                void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    int rc = glr_block_params_ok(s, p);
    if (rc) return rc;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(feat0); GLR_REQUIRE_PTR(feat1); GLR_REQUIRE_PTR(out);
    if (!sv) return GLRGTV_ERR_POINTER;
    float* need[9] = {sv->wT0, sv->wL0, sv->wT1, sv->wL1, sv->bA, sv->x1, sv->bB, sv->r1, sv->x2};
    for (int i = 0; i < 9; ++i) GLR_REQUIRE_PTR(need[i]);

    glrgtv_shape sc = *s;
    sc.H /= 2; sc.W /= 2;
    if ((rc = launch_weights(*s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, stream))) return rc;
    if ((rc = launch_weights(sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, stream))) return rc;

    BlockFwdArgs a;
    a.s = *s; a.p = *p;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1;
    a.y = nullptr; a.bB_in = nullptr; a.r1_in = nullptr; a.out1 = nullptr; a.out2 = nullptr;
    a.z = x; a.out0 = sv->bA;
    if ((rc = launch_stage<MODE_BA>(a, stream))) return rc;
    a.z = sv->bA; a.out0 = sv->x1;
    if ((rc = launch_stage<MODE_X1>(a, stream))) return rc;
    a.z = sv->x1; a.y = x; a.out0 = sv->x2; a.out1 = sv->bB; a.out2 = sv->r1;
    if ((rc = launch_stage<MODE_X2>(a, stream))) return rc;
    a.z = sv->x2; a.y = x; a.bB_in = sv->bB; a.r1_in = sv->r1; a.out0 = out; a.out1 = nullptr; a.out2 = nullptr;
    return launch_stage<MODE_X3>(a, stream);
}
