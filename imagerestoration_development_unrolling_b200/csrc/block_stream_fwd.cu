// block_stream_fwd.cu - register-streaming forward stages of LocalLowpassFilteringBlock / MixtureGTVGLR
// (V1X0:707-811, 985-988); same stage cut as block_fwd.cu (BA, X1, X2, X3), different execution model (stream.cuh).
//
// One CTA = one (batch, graph), `nch` of its channels, a band of rows.  Per channel there is a FINE walker (GL lanes,
// the full-resolution chain and the stage epilogue) and a COARSE walker (GL/2 lanes, the half-resolution chain on the
// 2x2 mean of the input).  Both run the SAME step code on the same register state; they differ in how a row is
// loaded (direct / pooled) and where the finished row goes (global memory / a two-slot shared-memory ring that hands
// the coarse term 0.25 * P^T[...] to the fine epilogue).  The coarse walker steps on even block steps only and runs
// DF = 8 steps ahead of the fine walker, which is exactly the depth of its pipeline in fine rows.
//
// Eligible shapes: W % 8 == 0 and W <= 256 (a walker is at most two warps wide); everything else takes the plane
// kernels of block_fwd.cu.
#include "stream.cuh"

enum { MODE_BA = 0, MODE_X1 = 1, MODE_X2 = 2, MODE_X3 = 3 };
enum { PL_Z = 0, PL_SA = 1, PL_SB = 2, PL_LA = 3, PL_OB = 4, PL_OT = 5, PL_COUNT = 6 };
#define STREAM_DF 8

struct StreamFwdArgs {
    glrgtv_shape s;
    glrgtv_block_params p;
    const float* z;      // stencil input: y | bA | x1 | x2
    const float* y;      // X2: y ; X3: x (skip path)
    const float* bB_in;  // X3
    const float* r1_in;  // X3
    const float *wT0, *wL0, *wT1, *wL1;
    const float *cT0, *cT1;   // symmetric GTV coefficients [B,G,2,H,W] / [B,G,2,H/2,W/2]
    float* out0;  // BA: bA | X1: x1 | X2: x2 | X3: out
    float* out1;  // X2: bB
    float* out2;  // X2: r1
    int nch;        // channels per CTA (divides F)
    int band_rows;  // fine rows per CTA (even)
    int n_bands;
};

// one resolution of one walker
struct Lvl {
    int H, W;
    const float* z;      // channel plane of the stage input AT THE FINE resolution (the coarse walker pools it)
    const float* wL;     // [4][H][W] of this (b, g)
    const float* wT;     // [4][H][W]
    const float* cT;     // [2][H][W]
    StatsTaps kL, kT;
    float aL, aT, Gam;
};

template <int MODE>
struct WState {
    Row z1, z2, z3;          // input rows t-1, t-2, t-3
    Row sA1, sA2, sB1, sB2;  // S rows t-2, t-3
    Row lA1, lA2, oB1, oB2;  // core rows t-3, t-4
    Row oT1, oT2;            // thresholded core (X2)
    Row cDp;                 // cD coefficients of core row t-3
};

template <int MODE, bool XW>
__global__ void __launch_bounds__(384) k_stream_fwd(StreamFwdArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool GLR = MODE != MODE_BA, THR = MODE == MODE_X2;
    constexpr int NRING = THR ? 2 : 1;
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int GL = XW ? 64 : (W <= 32 ? 8 : W <= 64 ? 16 : 32), GLc = GL / 2;
    const int nch = a.nch;
    const int NTF = (nch * GL + 31) & ~31;      // fine threads first, then coarse threads (roles never share a warp)
    const int tid = (int)threadIdx.x;
    const bool fine = tid < NTF;
    const int gl = fine ? GL : GLc;
    const int wk = fine ? tid / GL : (tid - NTF) / GLc;          // walker == channel slot
    const int lane = fine ? tid % GL : (tid - NTF) % GLc;
    const bool live = wk < nch;

    // block -> (band, channel chunk, graph, batch)
    int bid = (int)blockIdx.x;
    const int band = bid % a.n_bands; bid /= a.n_bands;
    const int chunks = F / nch;
    const int chunk = bid % chunks; bid /= chunks;
    const int g = bid % G, b = bid / G;
    const int f = chunk * nch + (live ? wk : 0), c = g * F + f;
    const int R0 = band * a.band_rows, R1 = R0 + a.band_rows < H ? R0 + a.band_rows : H;
    const size_t HW = (size_t)H * W, HWc = HW / 4;
    const size_t plane = (size_t)b * G + g;
    const size_t off = ((size_t)b * G * F + c) * HW;

    // shared memory: ring [nch][2 slots][NRING][2*GL]  |  mailboxes [2][nch][PL_COUNT][2]
    const int ringW = 2 * GL;
    float* ring = smem + (size_t)wk * 2 * NRING * ringW;
    float* mbox = smem + (size_t)nch * 2 * NRING * ringW;

    Lvl L;
    if (fine) {
        L.H = H; L.W = W;
        L.wL = a.wL0 + plane * 4 * HW; L.wT = a.wT0 + plane * 4 * HW; L.cT = a.cT0 + plane * 2 * HW;
        L.kT = glr_load_taps(a.p.gtv0.stats, c);
        L.kL = GLR ? glr_load_taps(a.p.glr0.stats, c) : L.kT;
        L.aT = expf(a.p.ro0[g]); L.aL = GLR ? expf(a.p.mu0[g]) : 0.f; L.Gam = THR ? expf(a.p.gamma0[g]) : 0.f;
    } else {
        L.H = H / 2; L.W = W / 2;
        L.wL = a.wL1 + plane * 4 * HWc; L.wT = a.wT1 + plane * 4 * HWc; L.cT = a.cT1 + plane * 2 * HWc;
        L.kT = glr_load_taps(a.p.gtv1.stats, c);
        L.kL = GLR ? glr_load_taps(a.p.glr1.stats, c) : L.kT;
        L.aT = expf(a.p.ro1[g]); L.aL = GLR ? expf(a.p.mu1[g]) : 0.f; L.Gam = THR ? expf(a.p.gamma1[g]) : 0.f;
    }
    L.z = a.z + off;
    const size_t LHW = (size_t)L.H * L.W;

    LaneCtx lc;
    lc.col0 = 4 * lane;
    lc.width = gl < 32 ? gl : 32;
    lc.active = live && lc.col0 < L.W;
    lc.first = lc.col0 == 0;
    lc.last = lc.col0 + 4 >= L.W;
    lc.seam_l = XW && live && fine && lane == 32 && lc.col0 < L.W;
    lc.seam_r = XW && live && fine && lane == 31 && lc.col0 + 4 < L.W;
    lc.mb_rd = lc.mb_wr = mbox;

    float alpha = 0.f, beta2 = 0.f, s0 = 0.f, s1 = 1.f;
    if (MODE == MODE_X1) alpha = a.p.alpha[0 * G + g];
    if (MODE == MODE_X2) alpha = a.p.alpha[1 * G + g];
    if (MODE == MODE_X3) {
        alpha = a.p.alpha[2 * G + g];
        beta2 = a.p.beta[2 * G + g];
        if (a.p.skip) { s0 = a.p.skip[0]; s1 = a.p.skip[1]; }
    }
    const bool has_skip = MODE == MODE_X3 && a.p.skip != nullptr;

    WState<MODE> st;
    st.z1 = st.z2 = st.z3 = st.sA1 = st.sA2 = st.sB1 = st.sB2 = st.lA1 = st.lA2 = st.oB1 = st.oB2 = st.oT1 = st.oT2 = st.cDp = row_zero();

    // the walker's row range at its own resolution and the first row it loads
    const int r0 = fine ? R0 : R0 / 2, r1 = fine ? R1 : R1 / 2;
    const int M = (R1 - R0) + 6 + STREAM_DF;

    for (int m = 0; m < M; ++m) {
        const bool stepping = fine ? (m >= STREAM_DF) : ((m & 1) == 0);
        if (stepping) {
            const int t = r0 - 3 + (fine ? m - STREAM_DF : (m >> 1));     // newest row of this step
            if (XW) {
                lc.mb_rd = mbox + (size_t)((m + 1) & 1) * nch * PL_COUNT * 2 + (size_t)wk * PL_COUNT * 2;
                lc.mb_wr = mbox + (size_t)(m & 1) * nch * PL_COUNT * 2 + (size_t)wk * PL_COUNT * 2;
            }
            // ---- load row t (fine: the row itself; coarse: the 2x2 mean of fine rows 2t, 2t+1)
            Row zn = row_zero();
            if (lc.active && t >= 0 && t < L.H) {
                if (fine) {
                    zn = row_ld(L.z + (size_t)t * W + lc.col0);
                } else {
                    const float* p = L.z + (size_t)(2 * t) * W + 2 * lc.col0;
                    const Row a0 = row_ld(p), a1 = row_ld(p + 4), b0 = row_ld(p + W), b1 = row_ld(p + W + 4);
                    zn.v[0] = 0.25f * (a0.v[0] + a0.v[1] + b0.v[0] + b0.v[1]);
                    zn.v[1] = 0.25f * (a0.v[2] + a0.v[3] + b0.v[2] + b0.v[3]);
                    zn.v[2] = 0.25f * (a1.v[0] + a1.v[1] + b1.v[0] + b1.v[1]);
                    zn.v[3] = 0.25f * (a1.v[2] + a1.v[3] + b1.v[2] + b1.v[3]);
                }
            }
            mb_post<XW>(zn, lc, PL_Z);
            // ---- S at row t-1
            Row sAn, sBn;
            {
                const int r = t - 1;
                const Row u = row_sel(r == 0, st.z1, st.z2), d = row_sel(r == L.H - 1, st.z1, zn);
                float l, rr;
                nb_lr<false, XW>(st.z1, l, rr, lc, PL_Z);
                sBn = w_S(L.kT, st.z1, u, d, l, rr);
                sAn = GLR ? w_S(L.kL, st.z1, u, d, l, rr) : sBn;
                if (GLR) mb_post<XW>(sAn, lc, PL_SA);
                mb_post<XW>(sBn, lc, PL_SB);
            }
            // ---- L and the GTV cores at row t-2 (zero rows outside the image)
            Row lAn = row_zero(), oBn = row_zero(), oTn = row_zero();
            {
                const int r = t - 2;
                const bool in = r >= 0 && r < L.H;
                const bool ld = in && lc.active;
                const bool top = r == 0, bot = r == L.H - 1;
                const float* wrow = (const float*)nullptr;
                if (GLR) {
                    const Row u = row_sel(top, st.sA1, st.sA2), d = row_sel(bot, st.sA1, sAn);
                    float l, rr;
                    nb_lr<false, XW>(st.sA1, l, rr, lc, PL_SA);
                    Row w[4];
                    wrow = L.wL + (size_t)(ld ? r : 0) * L.W + (lc.active ? lc.col0 : 0);
#pragma unroll
                    for (int e = 0; e < 4; ++e) w[e] = row_ld_if(ld, wrow + e * LHW);
                    lAn = row_sel(in, w_L(st.sA1, u, d, l, rr, w), lAn);
                }
                {
                    const Row u = row_sel(top, st.sB1, st.sB2), d = row_sel(bot, st.sB1, sBn);
                    float l, rr;
                    nb_lr<false, XW>(st.sB1, l, rr, lc, PL_SB);
                    const float* crow = L.cT + (size_t)(ld ? r : 0) * L.W + (lc.active ? lc.col0 : 0);
                    const Row cr = row_ld_if(ld, crow), cd = row_ld_if(ld, crow + LHW);
                    const float cr_left = (ld && lc.col0 > 0) ? crow[-1] : 0.f;
                    oBn = row_sel(in, w_core_lin(st.sB1, u, d, l, rr, cr, cr_left, cd, st.cDp), oBn);
                    st.cDp = cd;
                    if (THR) {
                        RawW rw;
                        if (ld) {
                            ld_raw_w(L.wT, L.H, L.W, r, lc.col0, rw);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) rw.own[e] = rw.in[e] = row_zero();
                        }
                        oTn = row_sel(in, w_core_thr(st.sB1, u, d, l, rr, rw, L.Gam), oTn);
                    }
                }
                if (GLR) mb_post<XW>(lAn, lc, PL_LA);
                mb_post<XW>(oBn, lc, PL_OB);
                if (THR) mb_post<XW>(oTn, lc, PL_OT);
            }
            // ---- St at row t-3 and the row's destination
            {
                const int r = t - 3;
                float l, rr;
                nb_lr<true, XW>(st.oB1, l, rr, lc, PL_OB);
                Row Az = w_St(L.kT, st.oB1, st.oB2, oBn, l, rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) Az.v[j] *= L.aT;
                if (GLR) {
                    nb_lr<true, XW>(st.lA1, l, rr, lc, PL_LA);
                    const Row gl_ = w_St(L.kL, st.lA1, st.lA2, lAn, l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) Az.v[j] += L.aL * gl_.v[j];
                }
                Row rT = row_zero();
                if (THR) {
                    nb_lr<true, XW>(st.oT1, l, rr, lc, PL_OT);
                    rT = w_St(L.kT, st.oT1, st.oT2, oTn, l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) rT.v[j] *= L.aT;
                }
                if (!fine) {
                    // coarse: hand 0.25 * (coarse term) to the fine epilogue of rows 2r, 2r+1
                    if (lc.active && r >= 0 && r < L.H) {
                        float* slot = ring + (size_t)(r & 1) * NRING * ringW + lc.col0;
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = 0.25f * Az.v[j];
                        st4(slot, v);
                        if (THR) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = 0.25f * rT.v[j];
                            st4(slot + ringW, v);
                        }
                    }
                } else if (lc.active && r >= R0 && r < R1) {
                    const float* slot = ring + (size_t)((r >> 1) & 1) * NRING * ringW + (lc.col0 >> 1);
                    const float2 cz = *reinterpret_cast<const float2*>(slot);
                    const Row zq = st.z3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) Az.v[j] += zq.v[j] + (j < 2 ? cz.x : cz.y);
                    if (THR) {
                        const float2 ct = *reinterpret_cast<const float2*>(slot + ringW);
#pragma unroll
                        for (int j = 0; j < 4; ++j) rT.v[j] += (j < 2 ? ct.x : ct.y);
                    }
                    const size_t gi = off + (size_t)r * W + lc.col0;
                    Row in0 = row_zero(), in1 = row_zero(), in2 = row_zero(), o0, o1, o2;
                    if (MODE == MODE_X2 || has_skip) in0 = row_ld(a.y + gi);
                    if (MODE == MODE_X3) { in1 = row_ld(a.bB_in + gi); in2 = row_ld(a.r1_in + gi); }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (MODE == MODE_BA) {
                            o0.v[j] = Az.v[j];  // y + R_lin(y)
                        } else if (MODE == MODE_X1) {
                            o0.v[j] = zq.v[j] + alpha * (zq.v[j] - Az.v[j]);
                        } else if (MODE == MODE_X2) {
                            const float bB = in0.v[j] + rT.v[j], r1v = bB - Az.v[j];
                            o1.v[j] = bB;
                            o2.v[j] = r1v;
                            o0.v[j] = zq.v[j] + alpha * r1v;
                        } else {
                            const float u2 = (in1.v[j] - Az.v[j]) + beta2 * in2.v[j];
                            const float x3 = zq.v[j] + alpha * u2;
                            o0.v[j] = has_skip ? s0 * in0.v[j] + s1 * x3 : x3;
                        }
                    }
                    st4(a.out0 + gi, o0.v);
                    if (MODE == MODE_X2) { st4(a.out1 + gi, o1.v); st4(a.out2 + gi, o2.v); }
                }
            }
            // ---- rotate the windows
            st.z3 = st.z2; st.z2 = st.z1; st.z1 = zn;
            st.sA2 = st.sA1; st.sA1 = sAn; st.sB2 = st.sB1; st.sB1 = sBn;
            st.lA2 = st.lA1; st.lA1 = lAn; st.oB2 = st.oB1; st.oB1 = oBn;
            if (THR) { st.oT2 = st.oT1; st.oT1 = oTn; }
        }
        if (XW || (m & 1)) __syncthreads();
    }
    (void)r1;
}

// symmetric GTV coefficients of one weight set: c[0] = wR^2 + wL[.,w+1]^2, c[1] = wD^2 + wU[h+1,.]^2 (0 for the missing neighbour)
__global__ void __launch_bounds__(256) k_gtv_coeffs(int planes, int H, int W, const float* __restrict__ w, float* __restrict__ c) {
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)planes * HW;
#ifdef GLRGTV_EMU
    for (size_t i = 0; i < total; ++i) {
#else
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
#endif
        const size_t pl = i / HW, o = i % HW;
        const int h = (int)(o / W), x = (int)(o % W);
        const float* wp = w + pl * 4 * HW;
        const float wr = wp[2 * HW + o], wd = wp[3 * HW + o];
        const float wl = x + 1 < W ? wp[1 * HW + o + 1] : 0.f;
        const float wu = h + 1 < H ? wp[o + W] : 0.f;
        c[pl * 2 * HW + o] = wr * wr + wl * wl;
        c[pl * 2 * HW + HW + o] = wd * wd + wu * wu;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int g_glr_block_path = 0;   // 0 auto, 1 plane kernels only, 2 streaming kernels (error when the shape is not eligible)
extern "C" int glrgtv_set_block_path(int mode) {
    if (mode < 0 || mode > 2) return GLRGTV_ERR_UNSUPPORTED;
    g_glr_block_path = mode;
    return GLRGTV_OK;
}

int glr_stream_eligible(const glrgtv_shape* s) {
    return s->W % 8 == 0 && s->W <= 256 && s->H % 2 == 0 && s->H >= 2;
}

struct StreamPlan {
    int GL, nch, threads, band_rows, n_bands;
    size_t smem;
};
static StreamPlan stream_plan(const glrgtv_shape& s, int nring) {
    StreamPlan p;
    p.GL = s.W > 128 ? 64 : s.W > 64 ? 32 : s.W > 32 ? 16 : 8;
    p.nch = 1;
    for (int n = 1; n <= s.F; ++n) {
        if (s.F % n) continue;
        const int thr = ((n * p.GL + 31) & ~31) + ((n * p.GL / 2 + 31) & ~31);
        if (thr <= 320) p.nch = n;
    }
    p.threads = ((p.nch * p.GL + 31) & ~31) + ((p.nch * p.GL / 2 + 31) & ~31);
    // whole-height bands unless the grid would leave SMs idle
    const long ctas = (long)s.B * s.G * (s.F / p.nch);
    int bands = 1;
    while (ctas * bands < 296 && s.H / (bands * 2) >= 32) bands *= 2;
    p.band_rows = ((s.H + bands - 1) / bands + 1) & ~1;
    p.n_bands = (s.H + p.band_rows - 1) / p.band_rows;
    p.smem = ((size_t)p.nch * 2 * nring * 2 * p.GL + (size_t)2 * p.nch * PL_COUNT * 2 + 8) * sizeof(float);
    return p;
}

template <int MODE>
static int launch_stream_stage(StreamFwdArgs a, void* stream) {
    const glrgtv_shape& s = a.s;
    const StreamPlan p = stream_plan(s, MODE == MODE_X2 ? 2 : 1);
    a.nch = p.nch; a.band_rows = p.band_rows; a.n_bands = p.n_bands;
    const long blocks = (long)s.B * s.G * (s.F / p.nch) * p.n_bands;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_BA + MODE, stream);
    if (p.GL == 64) GLR_LAUNCH_FIBERS((k_stream_fwd<MODE, true>), dim3((unsigned)blocks), p.threads, p.smem, stream, a);
    else GLR_LAUNCH_FIBERS((k_stream_fwd<MODE, false>), dim3((unsigned)blocks), p.threads, p.smem, stream, a);
    GLR_PROF_END(GLRGTV_SLOT_FWD_BA + MODE, stream);
    return GLR_CHECK_LAUNCH();
}

int glr_launch_gtv_coeffs(const glrgtv_shape& s, const float* w, float* c, void* stream) {
    const long planes = (long)s.B * s.G;
    const size_t total = (size_t)planes * s.H * s.W;
#ifdef GLRGTV_EMU
    const unsigned blocks = 1;
#else
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
#endif
    GLR_LAUNCH(k_gtv_coeffs, dim3(blocks ? blocks : 1), 256, 0, stream, (int)planes, s.H, s.W, w, c);
    return GLR_CHECK_LAUNCH();
}

// the four solver stages on the streaming kernels; weights and coefficient planes are already in `sv`
int glr_stream_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x, float* out,
                         const glrgtv_block_saved* sv, void* stream) {
    int rc;
    StreamFwdArgs a;
    a.s = *s; a.p = *p;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1; a.cT0 = sv->cT0; a.cT1 = sv->cT1;
    a.y = nullptr; a.bB_in = nullptr; a.r1_in = nullptr; a.out1 = nullptr; a.out2 = nullptr;
    a.nch = 1; a.band_rows = s->H; a.n_bands = 1;
    a.z = x; a.out0 = sv->bA;
    if ((rc = launch_stream_stage<MODE_BA>(a, stream))) return rc;
    a.z = sv->bA; a.out0 = sv->x1;
    if ((rc = launch_stream_stage<MODE_X1>(a, stream))) return rc;
    a.z = sv->x1; a.y = x; a.out0 = sv->x2; a.out1 = sv->bB; a.out2 = sv->r1;
    if ((rc = launch_stream_stage<MODE_X2>(a, stream))) return rc;
    a.z = sv->x2; a.y = x; a.bB_in = sv->bB; a.r1_in = sv->r1; a.out0 = out; a.out1 = nullptr; a.out2 = nullptr;
    return launch_stream_stage<MODE_X3>(a, stream);
}
