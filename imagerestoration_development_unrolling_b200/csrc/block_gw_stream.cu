// block_gw_stream.cu - edge-weight gradients of one backward stage as streaming row walkers (stream.cuh).
//
// Same quantities as block_gw.cu (which remains the path for planes the walkers do not cover):
//   L :      gw_e[p] -= sum_f  h_L[p] * s_L[n_e(p)]                                   s_K = S_K z, h_K = a_K S0_K g
//   T lin :  gw_e[p] += sum_f  2 w_e (h_T[p] - h_T[n_e]) (s_T[p] - s_T[n_e])
//   T thr :  gw_e[p] += sum_f  D2 phi(w d) + D2 w phi'(w d) d                          (stage X2, upstream gB -> h2)
// They need only the first-level stencils, so a walker here keeps four (six) 3-row windows: z, g, s, h (, g2, h2).
// One CTA owns a (batch, graph) at one resolution and ALL F channels of the graph: per channel a T walker (gwT) and,
// where the stage has a GLR part, an L walker (gwL).  Every step each walker posts the four gradient rows of its
// channel to a double-buffered shared-memory board; one step later all threads sum the board over the channels and
// read-modify-write the rows of gw - the reduction over channels never leaves the SM.  Operands arrive through
// cp.async rings two steps ahead (the coarse resolution pools two fine rows while reading the ring).
#include "stream.cuh"
#include "stream_bwd.cuh"

enum { GS_S = 0, GS_H = 1, GS_H2 = 2, GS_COUNT = 3 };
#define GS_ZR 4          // ring depth in rows of this resolution (coarse: x2 fine rows each)
#define GS_WR 4
#define GS_PD 2

template <int MODE, bool COARSE>
struct GsSmem {
    static constexpr bool HAS_L = MODE != BW_BA, X2 = MODE == BW_X2A;
    static constexpr int NSRC = X2 ? 2 : 1, NK = HAS_L ? 2 : 1, NP = HAS_L ? 8 : 4, RPS = COARSE ? 2 : 1;   // ring rows per step
    int Wr, Wl, F;     // ring row length (fine columns, padded), level row length (padded), channels
    __host__ __device__ size_t zring() const { return 0; }                                                    // [F][ZR*RPS][Wr]
    __host__ __device__ size_t sring() const { return zring() + (size_t)F * GS_ZR * RPS * Wr; }              // [NSRC][F][ZR*RPS][Wr]
    __host__ __device__ size_t wring() const { return sring() + (size_t)NSRC * F * GS_ZR * RPS * Wr; }       // [4][WR][Wl]
    __host__ __device__ size_t board() const { return wring() + (size_t)4 * GS_WR * Wl; }                    // [2][F][NP][Wl]
    __host__ __device__ size_t mbox() const { return board() + (size_t)2 * F * NP * Wl; }                    // [2][F][NK][GS_COUNT][2]
    __host__ __device__ size_t total() const { return mbox() + (size_t)2 * F * NK * GS_COUNT * 2 + 8; }
};

template <int MODE, bool XW, bool COARSE>
__global__ void __launch_bounds__(768, 1) k_gw_stream(GwArgs a) {
    GLR_SMEM_DECL(smem);
    using SM = GsSmem<MODE, COARSE>;
    constexpr bool HAS_L = SM::HAS_L, X2 = SM::X2;
    constexpr int NSRC = SM::NSRC, NK = SM::NK, NP = SM::NP, RPS = SM::RPS, ZR = GS_ZR, PD = GS_PD;
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int LH = COARSE ? H / 2 : H, LW = COARSE ? W / 2 : W, LHW = LH * LW;
    const int GL = XW ? 64 : (LW <= 32 ? 8 : LW <= 64 ? 16 : 32);          // lanes of one walker at THIS resolution
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    const int NTK = (F * GL + 31) & ~31;                                     // threads of one kind (T first, then L)
    const bool ISL = HAS_L && tid >= NTK;
    const int tk = ISL ? tid - NTK : tid;
    const int wk = tk / GL, lane = tk % GL, nwk = NTK / GL;
    const bool live = wk < F;
    const int wkc = live ? wk : 0;
    const int g = (int)blockIdx.x % G, b = (int)blockIdx.x / G;
    const int c = g * F + wkc;
    const size_t HW = (size_t)H * W, off = ((size_t)b * G * F + c) * HW, plane = (size_t)b * G + g;

    SM lay; lay.Wl = 4 * GL; lay.Wr = COARSE ? 8 * GL : 4 * GL; lay.F = F;
    const int Wl = lay.Wl, Wr = lay.Wr;
    {
        const int n4 = (int)(lay.total() / 4);
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += NT) st4(smem + 4 * i, z4);
        __syncthreads();
    }
    const float* zring = smem + lay.zring() + (size_t)wkc * ZR * RPS * Wr;
    const float* s0ring = smem + lay.sring() + (size_t)wkc * ZR * RPS * Wr;
    const float* s1ring = s0ring + (size_t)F * ZR * RPS * Wr;
    const float* wring = smem + lay.wring();
    float* board = smem + lay.board();
    float* mbox = smem + lay.mbox();
    const smem_addr_t sbase = smem_addr(smem);

    // ---- scalars: the stage's upstream combination(s) and this walker's module
    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const float s1 = a.p.skip ? a.p.skip[1] : 1.f, c23 = al2 * s1;
    float ca, cb = 0.f, ca2 = 0.f, cb2 = 0.f;
    if (MODE == BW_X3) ca = -c23;
    else if (X2) { ca = -be2 * c23; cb = -al1; ca2 = c23 + be2 * c23; cb2 = al1; }
    else if (MODE == BW_X1) ca = -al0;
    else ca = 1.f;
    const glrgtv_opparams& op = ISL ? (COARSE ? a.p.glr1 : a.p.glr0) : (COARSE ? a.p.gtv1 : a.p.gtv0);
    const StatsTaps kK = glr_load_taps(op.stats, c);
    const float aK = expf(ISL ? (COARSE ? a.p.mu1[g] : a.p.mu0[g]) : (COARSE ? a.p.ro1[g] : a.p.ro0[g]));
    const float Gam = X2 ? expf(COARSE ? a.p.gamma1[g] : a.p.gamma0[g]) : 0.f;
    const float* wT = (COARSE ? a.wT1 : a.wT0) + plane * 4 * LHW;
    float* gwT = (COARSE ? a.gwT1 : a.gwT0) + plane * 4 * LHW;
    float* gwL = (COARSE ? a.gwL1 : a.gwL0) + plane * 4 * LHW;

    LaneCtx lc;
    lc.col0 = 4 * lane;
    lc.width = GL < 32 ? GL : 32;
    lc.active = live && lc.col0 < LW;
    lc.first = lc.col0 == 0;
    lc.last = lc.col0 + 4 >= LW;
    lc.seam_l = XW && live && lane == 32 && lc.col0 < LW;
    lc.seam_r = XW && live && lane == 31 && lc.col0 + 4 < LW;
    lc.mb_rd = lc.mb_wr = mbox;

    // ---- loader: T walkers stage their channel's operand rows (own columns), walkers 0..3 of kind T the weight planes
    const int cols = COARSE ? 2 * lc.col0 : lc.col0;                  // first fine column of this lane
    const smem_addr_t zdst = smem_advance(sbase, (int)lay.zring() + wkc * ZR * RPS * Wr + cols);
    const smem_addr_t s0dst = smem_advance(sbase, (int)lay.sring() + wkc * ZR * RPS * Wr + cols);
    const smem_addr_t s1dst = smem_advance(s0dst, F * ZR * RPS * Wr);
    auto issue = [&](int mt) {
        const int t = mt - 0;                       // row of this resolution a step loads: t = step index (rows 0 .. LH-1)
        if (t < 0 || t >= LH) return;
        if (!ISL && lc.active) {
#pragma unroll
            for (int k = 0; k < RPS; ++k) {
                const int rho = RPS * t + k, so = ((t & (ZR - 1)) * RPS + k) * Wr;
                const size_t go = off + (size_t)rho * W + cols;
#pragma unroll
                for (int q = 0; q < RPS; ++q) {
                    cp_async16_s(smem_advance(zdst, so + 4 * q), a.z + go + 4 * q);
                    cp_async16_s(smem_advance(s0dst, so + 4 * q), a.src0 + go + 4 * q);
                    if (NSRC == 2) cp_async16_s(smem_advance(s1dst, so + 4 * q), a.src1 + go + 4 * q);
                }
            }
        }
    };
    // raw GTV weights of the row whose gradients step `mt` computes: row mt - 2
    auto issue_w = [&](int mt) {
        const int rw = mt - 2;
        if (rw < 0 || rw >= LH) return;
        if (!ISL && lc.col0 < LW)
            for (int e = wk; e < 4; e += nwk)
                cp_async16_s(smem_advance(sbase, (int)lay.wring() + (e * GS_WR + (rw & (GS_WR - 1))) * Wl + lc.col0),
                             wT + (size_t)e * LHW + (size_t)rw * LW + lc.col0);
    };

    Row z[3], gq[3], g2[3], s[3], h[3], h2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) z[k] = gq[k] = g2[k] = s[k] = h[k] = h2[k] = row_zero();
    float gam = 0.f;
    constexpr int N = 2, C = 1, U = 0;
    const int M = LH + 3;                      // rows 0..LH-1 are loaded at steps 0..LH-1; gradients of row r at step r+2; board at r+3

#pragma unroll
    for (int k = 0; k < PD; ++k) { issue(k); issue_w(k); cp_async_commit(); }

#pragma unroll 1
    for (int m = 0; m < M; ++m) {
        cp_async_wait_pending<PD - 1>();
        __syncthreads();
        issue(m + PD);
        issue_w(m + PD);
        cp_async_commit();
        const int t = m;
        if (XW) {
            lc.mb_rd = mbox + (((m + 1) & 1) * F * NK + wkc * NK + (ISL ? 1 : 0)) * GS_COUNT * 2;
            lc.mb_wr = mbox + ((m & 1) * F * NK + wkc * NK + (ISL ? 1 : 0)) * GS_COUNT * 2;
        }
        // ---- reduction of the rows posted one step ago (gradient row t-3): sum the board over the channels, update gw
        {
            const int r = t - 3;
            if (r >= 0 && r < LH) {
                const float* bd = board + (size_t)((m + 1) & 1) * F * NP * Wl;
                const int Q = LW / 4;
                for (int i = tid; i < NP * Q; i += NT) {
                    const int p = i / Q, q = i - p * Q;
                    Row acc = row_ld(bd + p * Wl + 4 * q);
                    for (int f = 1; f < F; ++f) {
                        const Row v = row_ld(bd + ((size_t)f * NP + p) * Wl + 4 * q);
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc.v[j] += v.v[j];
                    }
                    // planes 0..3: gwT edges; 4..7: gwL edges
                    float* dst = (p < 4 ? gwT + (size_t)p * LHW : gwL + (size_t)(p - 4) * LHW) + (size_t)r * LW + 4 * q;
                    if (!a.assign) {
                        const Row o = row_ld(dst);
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc.v[j] += o.v[j];
                    }
                    st4(dst, acc.v);
                }
            }
        }
        // ---- rows t of z (clamp-extended at production) and of the upstream(s) (zero outside)
        if (t < LH) {
            const int so = (t & (ZR - 1)) * RPS * Wr + cols;
            if (!COARSE) {
                z[N] = row_ld(zring + so);
                const Row q0 = row_ld(s0ring + so);
                Row q1 = row_zero();
                if (NSRC == 2) q1 = row_ld(s1ring + so);
#pragma unroll
                for (int j = 0; j < 4; ++j) { gq[N].v[j] = ca * q0.v[j] + cb * q1.v[j]; if (X2) g2[N].v[j] = ca2 * q0.v[j] + cb2 * q1.v[j]; }
            } else {
                auto pool = [&](const float* ring, Row& dst) {
                    const float* p0 = ring + so;
                    const Row a0 = row_ld(p0), a1 = row_ld(p0 + 4), b0 = row_ld(p0 + Wr), b1 = row_ld(p0 + Wr + 4);
                    dst.v[0] = 0.25f * (a0.v[0] + a0.v[1] + b0.v[0] + b0.v[1]);
                    dst.v[1] = 0.25f * (a0.v[2] + a0.v[3] + b0.v[2] + b0.v[3]);
                    dst.v[2] = 0.25f * (a1.v[0] + a1.v[1] + b1.v[0] + b1.v[1]);
                    dst.v[3] = 0.25f * (a1.v[2] + a1.v[3] + b1.v[2] + b1.v[3]);
                };
                pool(zring, z[N]);
                Row q0, q1 = row_zero();
                pool(s0ring, q0);
                if (NSRC == 2) pool(s1ring, q1);
#pragma unroll
                for (int j = 0; j < 4; ++j) { gq[N].v[j] = ca * q0.v[j] + cb * q1.v[j]; if (X2) g2[N].v[j] = ca2 * q0.v[j] + cb2 * q1.v[j]; }
            }
            if (t == 0) z[C] = z[N];
        } else {
            z[N] = z[C];
            gq[N] = row_zero();
            if (X2) g2[N] = row_zero();
        }
        // ---- first-level stencils at row t-1
        {
            const int r = t - 1;
            if (r >= 0 && r < LH) {
                float l, rr;
                nb_lr<false, false>(z[C], l, rr, lc, 0);
                if (XW) {       // a 64-lane walker: the seam scalars of z and g come from the ring rows
                    const float* pz = zring + (r & (ZR - 1)) * RPS * Wr + cols;
                    if (lc.seam_l) l = pz[-1];
                    if (lc.seam_r) rr = pz[4];
                }
                s[N] = w_S(kK, z[C], z[U], z[N], l, rr);
                auto g_lr = [&](const Row& gc, float ka, float kb, float& gl_, float& gr_) {
                    nb_lr<true, false>(gc, gl_, gr_, lc, 0);
                    if (XW) {
                        const int so = (r & (ZR - 1)) * RPS * Wr + cols;
                        if (lc.seam_l) gl_ = ka * s0ring[so - 1] + (NSRC == 2 ? kb * s1ring[so - 1] : 0.f);
                        if (lc.seam_r) gr_ = ka * s0ring[so + 4] + (NSRC == 2 ? kb * s1ring[so + 4] : 0.f);
                    }
                };
                g_lr(gq[C], ca, cb, l, rr);
                h[N] = w_S(kK, gq[C], gq[U], gq[N], l, rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[N].v[j] *= aK;
                if (X2 && !ISL) {
                    g_lr(g2[C], ca2, cb2, l, rr);
                    h2[N] = w_S(kK, g2[C], g2[U], g2[N], l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h2[N].v[j] *= aK;
                }
                if (r == 0) { s[C] = s[N]; if (!ISL) { h[C] = h[N]; if (X2) h2[C] = h2[N]; } }
            } else if (r >= LH) {
                s[N] = s[C];
                h[N] = ISL ? row_zero() : h[C];
                if (X2) h2[N] = h2[C];
            }
            mb_post<XW>(s[N], lc, GS_S);
            mb_post<XW>(h[N], lc, GS_H);
            if (X2) mb_post<XW>(h2[N], lc, GS_H2);
        }
        // ---- gradients of row t-2, posted to the board
        {
            const int r = t - 2;
            if (r >= 0 && r < LH) {
                float sl, sr;
                nb_lr<false, XW>(s[C], sl, sr, lc, GS_S);
                Row acc[4];
                if (ISL) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float hv = h[C].v[j];
                        acc[0].v[j] = -hv * s[U].v[j];
                        acc[1].v[j] = -hv * rowL(s[C], sl, j);
                        acc[2].v[j] = -hv * rowR(s[C], sr, j);
                        acc[3].v[j] = -hv * s[N].v[j];
                    }
                } else {
                    float hl, hr, h2l = 0.f, h2r = 0.f;
                    nb_lr<false, XW>(h[C], hl, hr, lc, GS_H);
                    if (X2) nb_lr<false, XW>(h2[C], h2l, h2r, lc, GS_H2);
                    Row we[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) we[e] = row_ld(wring + (e * GS_WR + (r & (GS_WR - 1))) * Wl + lc.col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d[4] = {s[C].v[j] - s[U].v[j], s[C].v[j] - rowL(s[C], sl, j), s[C].v[j] - rowR(s[C], sr, j), s[C].v[j] - s[N].v[j]};
                        const float D[4] = {h[C].v[j] - h[U].v[j], h[C].v[j] - rowL(h[C], hl, j), h[C].v[j] - rowR(h[C], hr, j), h[C].v[j] - h[N].v[j]};
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e].v[j] = 2.f * we[e].v[j] * D[e] * d[e];
                        if (X2) {
                            const float D2[4] = {h2[C].v[j] - h2[U].v[j], h2[C].v[j] - rowL(h2[C], h2l, j), h2[C].v[j] - rowR(h2[C], h2r, j),
                                                 h2[C].v[j] - h2[N].v[j]};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float w = we[e].v[j], tt = w * d[e];
                                acc[e].v[j] += D2[e] * glr_phi(tt, Gam) + D2[e] * w * glr_dphi(tt, Gam) * d[e];
                                if (lc.active && fabsf(tt) > Gam) gam += D2[e] * w * (tt > 0.f ? -2.f : 2.f);
                            }
                        }
                    }
                }
                if (lc.active) {
                    float* bd = board + ((size_t)(m & 1) * F + wk) * NP * Wl + (ISL ? 4 : 0) * Wl + lc.col0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) st4(bd + e * Wl, acc[e].v);
                }
            }
        }
        // ---- rotate the windows
        z[U] = z[C]; z[C] = z[N]; gq[U] = gq[C]; gq[C] = gq[N]; s[U] = s[C]; s[C] = s[N]; h[U] = h[C]; h[C] = h[N];
        if (X2) { g2[U] = g2[C]; g2[C] = g2[N]; h2[U] = h2[C]; h2[C] = h2[N]; }
    }
    cp_async_wait_all();
    if (X2) {
        float v = (!ISL && live) ? gam : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v != 0.f) atomicAdd((COARSE ? a.ggamma1 : a.ggamma0) + g, v * Gam);
    }
}

extern unsigned long long g_glr_stream_launches;

// planes this kernel covers: the level's width is a multiple of 4, at most 256, and the CTA fits 768 threads
static bool gs_ok(const glrgtv_shape& s, bool coarse, bool has_l, size_t* smem_out, int* threads_out, bool* xw_out) {
    const int LW = coarse ? s.W / 2 : s.W;
    if (LW % 4 || LW > 256 || LW < 4) return false;
    const int GL = LW > 128 ? 64 : LW <= 32 ? 8 : LW <= 64 ? 16 : 32;
    const int threads = (has_l ? 2 : 1) * ((s.F * GL + 31) & ~31);
    if (threads > 768) return false;
    *threads_out = threads; *xw_out = GL == 64;
    const size_t Wl = 4 * GL, Wr = coarse ? 8 * GL : 4 * GL, RPS = coarse ? 2 : 1;
    const size_t nsrc_max = 2, NP = has_l ? 8 : 4;
    const size_t fl = (1 + nsrc_max) * s.F * GS_ZR * RPS * Wr + 4 * GS_WR * Wl + 2 * s.F * NP * Wl + 2 * s.F * 2 * GS_COUNT * 2 + 8;
    *smem_out = fl * sizeof(float);
    return *smem_out <= 220 * 1024;
}

template <int MODE, bool XW, bool COARSE>
static int launch_gs(const GwArgs& a, int threads, void* stream) {
    GsSmem<MODE, COARSE> lay;
    const int LW = COARSE ? a.s.W / 2 : a.s.W;
    const int GL = XW ? 64 : (LW <= 32 ? 8 : LW <= 64 ? 16 : 32);
    lay.Wl = 4 * GL; lay.Wr = COARSE ? 8 * GL : 4 * GL; lay.F = a.s.F;
    const size_t smem = lay.total() * sizeof(float);
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_gw_stream<MODE, XW, COARSE>, smem, optin)) return rc_;
#endif
    ++g_glr_stream_launches;
    GLR_LAUNCH_FIBERS((k_gw_stream<MODE, XW, COARSE>), dim3((unsigned)(a.s.B * a.s.G)), threads, smem, stream, a);
    return GLRGTV_OK;
}

// returns GLRGTV_ERR_UNSUPPORTED when a level of this shape is outside the walkers' range (the caller then uses block_gw.cu)
template <int MODE>
int glr_gw_stream_stage(const GwArgs& a, int slot, void* stream) {
    constexpr bool HAS_L = MODE != BW_BA;
    size_t smem[2]; int threads[2]; bool xw[2];
    for (int lvl = 0; lvl < 2; ++lvl)
        if (!gs_ok(a.s, lvl == 1, HAS_L, &smem[lvl], &threads[lvl], &xw[lvl])) return GLRGTV_ERR_UNSUPPORTED;
    for (int lvl = 0; lvl < 2; ++lvl) {
        GLR_PROF_BEGIN(slot, stream);
        int rc;
        if (lvl == 0) rc = xw[0] ? launch_gs<MODE, true, false>(a, threads[0], stream) : launch_gs<MODE, false, false>(a, threads[0], stream);
        else rc = xw[1] ? launch_gs<MODE, true, true>(a, threads[1], stream) : launch_gs<MODE, false, true>(a, threads[1], stream);
        GLR_PROF_END(slot, stream);
        if (rc) return rc;
        if ((rc = GLR_CHECK_LAUNCH())) return rc;
    }
    return GLRGTV_OK;
}
template int glr_gw_stream_stage<BW_X3>(const GwArgs&, int, void*);
template int glr_gw_stream_stage<BW_X2A>(const GwArgs&, int, void*);
template int glr_gw_stream_stage<BW_X1>(const GwArgs&, int, void*);
template int glr_gw_stream_stage<BW_BA>(const GwArgs&, int, void*);
