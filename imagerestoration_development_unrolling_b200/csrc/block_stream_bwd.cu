// block_stream_bwd.cu - register-streaming backward stages of LocalLowpassFilteringBlock / MixtureGTVGLR
// (SURVEY Appendix B.9), the reverse sweep of block_stream_fwd.cu.
//
// A backward stage needs, per channel and resolution, the forward chain (recomputed) AND its adjoint chain - twelve
// row windows, more than a thread can hold.  The chains of the two operator families never mix before the final sum,
// so they get SEPARATE walkers (stream.cuh):
//     T walker:  z -S_T-> s -core-> o -St_T-> fwd_T          g -S0_T-> h -core'-> gs -S'_T-> V_T
//     L walker:  z -S_L-> s -L----> o -St_L-> fwd_L          g -S0_L-> h -L'----> gs -S'_L-> V_L
// (g = the stage's upstream gradient, a pointwise combination of one or two tensors; primes are VJPs).  Each walker
// holds six 3-row windows, accumulates the gradients of its own module's stats_kernel_p* and of its mu / ro in
// registers, and posts the rows V and fwd of every finished image row to shared memory; one step later the fine T
// walker ("finisher") adds them up with the coarse walkers' rows and the pointwise terms and writes the stage output.
// The thresholded part of stage X2 is its own launch (BW_X2B, T walkers only) that accumulates onto X2A's output.
//
// Edge-weight gradients are NOT computed here: they only need the first-level stencils (s, h) and are a sum over a
// graph's channels, which live in different CTAs here; block_gw.cu does them in one pass per stage.
//
// Planes wider than a 64-lane walker are cut into column strips of BW_STRIP valid columns with 8 halo columns per side
// (the reach of one stage through the half-resolution branch); halo lanes compute throw-away values and are excluded
// from every sum and store.
//
//   BW_X3 :  g = -a2 s1 gout                         z = x2   -> gx2
//   BW_X2A:  g = -(b2 a2 s1 gout + a1 gx2)           z = x1   -> gx1 (A part)
//   BW_X2B:  g =  (1+b2) a2 s1 gout + a1 gx2         z = x1   -> gx1 += (thresholded R part)
//   BW_X1 :  g = -a0 gx1                             z = bA   -> gbA
//   BW_BA :  g = gbA                                 z = y    -> gx (direct path)
#include "stream.cuh"
#include "stream_bwd.cuh"

enum { BP_G = 0, BP_S = 1, BP_H = 2, BP_O = 3, BP_GS = 4, BP_COUNT = 5 };
#define BW_DF 6      // block steps the coarse walkers run ahead of the fine ones (the depth of their pipeline in fine rows)
#define BW_PD 2
#define BW_ZR 12
#define BW_OPR 4
#define BW_WR 4
#define BW_MAXT 192
#define BW_MAXJ 4

template <int MODE>
struct BwSmem {
    static constexpr bool HAS_L = MODE == BW_X3 || MODE == BW_X2A || MODE == BW_X1, THR = MODE == BW_X2B;
    static constexpr int NSRC = (MODE == BW_X2A || MODE == BW_X2B) ? 2 : 1;
    static constexpr int NOP = MODE == BW_X3 ? 3 : MODE == BW_BA ? 2 : MODE == BW_X1 ? 0 : 1;
    static constexpr int NPT = THR ? 4 : 2, NPL = NPT + (HAS_L ? 4 : 0), NK = HAS_L ? 2 : 1;
    int Wp, nch;
    __host__ __device__ size_t zring() const { return 0; }
    __host__ __device__ size_t sring() const { return zring() + (size_t)nch * BW_ZR * Wp; }                  // [NSRC][nch][ZR][Wp]
    __host__ __device__ size_t opring() const { return sring() + (size_t)NSRC * nch * BW_ZR * Wp; }          // [nch][NOP][OPR][Wp]
    __host__ __device__ size_t w0ring() const { return opring() + (size_t)nch * NOP * BW_OPR * Wp; }         // [NPL][WR][Wp]
    __host__ __device__ size_t w1ring() const { return w0ring() + (size_t)NPL * BW_WR * Wp; }                // [NPL][WR][Wp/2]
    __host__ __device__ size_t xring() const { return w1ring() + (size_t)NPL * BW_WR * (Wp / 2); }           // [nch][2][NK][2][Wp]
    __host__ __device__ size_t cring() const { return xring() + (size_t)nch * 2 * NK * 2 * Wp; }             // [nch][4][NK][2][Wp/2]
    __host__ __device__ size_t mbox() const { return cring() + (size_t)nch * 4 * NK * 2 * (Wp / 2); }        // [2][nch][NK][BP_COUNT][2]
    __host__ __device__ size_t total() const { return mbox() + (size_t)2 * nch * NK * BP_COUNT * 2 + 8; }
};

// raw weights of row r from a 4-plane ring whose plane U (edge 0) is staged one row ahead: plane U of row r and plane
// D of row r-1 are carried in registers from the previous step.  `ring` points at this lane's quad of plane 0.
__device__ __forceinline__ void ring_raw_w(const float* ring, int wpitch, int r, int LH, bool first, bool last, Row& wU_next,
                                           Row& wD_prev, RawW& o) {
    auto prow = [&](int e, int row) { return ring + (e * BW_WR + (row & (BW_WR - 1))) * wpitch; };
    o.own[0] = wU_next;
    o.in[0] = wD_prev;
    const float* pL = prow(1, r);
    const float* pR = prow(2, r);
    o.own[1] = row_ld(pL);
    o.own[2] = row_ld(pR);
    o.own[3] = row_ld(prow(3, r));
    o.in[3] = r + 1 < LH ? row_ld(prow(0, r + 1)) : row_zero();   // edge U of the lower neighbour
    const float wr_m1 = first ? 0.f : pR[-1];                      // edge R of the left neighbour
    const float wl_p4 = last ? 0.f : pL[4];                        // edge L of the right neighbour
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.in[1].v[j] = j == 0 ? wr_m1 : o.own[2].v[j - 1];
        o.in[2].v[j] = j == 3 ? wl_p4 : o.own[1].v[j + 1];
    }
    wU_next = o.in[3];
    wD_prev = o.own[3];
}

// VJP of L wrt its input (tile.cuh q_L_adj in row form): h zero-extended
__device__ __forceinline__ Row w_L_adj(const Row& c, const Row& u, const Row& d, float l, float r, const RawW& w, bool top,
                                       bool bot, bool first, bool last) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = c.v[j] - (w.in[0].v[j] * u.v[j] + w.in[1].v[j] * rowL(c, l, j) + w.in[2].v[j] * rowR(c, r, j) + w.in[3].v[j] * d.v[j]);
        float self = 0.f;
        if (top) self += w.own[0].v[j];
        if (first && j == 0) self += w.own[1].v[j];
        if (last && j == 3) self += w.own[2].v[j];
        if (bot) self += w.own[3].v[j];
        o.v[j] = v - self * c.v[j];
    }
    return o;
}
// VJP wrt s of the thresholded GTV core (tile.cuh q_gtv_raw_adj<false, true> in row form): hB, s clamp-extended
__device__ __forceinline__ Row w_core_thr_adj(const Row& hc, const Row& hu, const Row& hd, float hl, float hr, const Row& sc,
                                              const Row& su, const Row& sd, float sl, float sr, const RawW& w, float Gam) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float D[4] = {hc.v[j] - hu.v[j], hc.v[j] - rowL(hc, hl, j), hc.v[j] - rowR(hc, hr, j), hc.v[j] - hd.v[j]};
        const float d[4] = {sc.v[j] - su.v[j], sc.v[j] - rowL(sc, sl, j), sc.v[j] - rowR(sc, sr, j), sc.v[j] - sd.v[j]};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float wa = w.own[e].v[j], wb = w.in[e].v[j];
            acc += D[e] * (wa * wa * glr_dphi(wa * d[e], Gam) + wb * wb * glr_dphi(wb * d[e], Gam));
        }
        o.v[j] = acc;
    }
    return o;
}

#define BW_STRIP 240     // valid output columns of one strip of a plane wider than a 64-lane walker (8 halo columns per side)
struct BwCta {
    int H, W, F, G, nch, b, g, f0, R0, R1, K0, K1, M, Wp;
    int x0, v0, v1;      // first column of the walker's window; valid output columns [v0, v1)
    size_t HW;
};
__device__ __forceinline__ BwCta bw_cta(const StreamBwdArgs& a, int GL) {
    BwCta c;
    c.H = a.s.H; c.W = a.s.W; c.F = a.s.F; c.G = a.s.G; c.nch = a.nch;
    int bid = (int)blockIdx.x;
    const int strip = bid % a.n_strips; bid /= a.n_strips;
    const int band = bid % a.n_bands; bid /= a.n_bands;
    const int chunks = c.F / c.nch;
    const int chunk = bid % chunks; bid /= chunks;
    c.g = bid % c.G; c.b = bid / c.G;
    c.f0 = chunk * c.nch;
    c.R0 = band * a.band_rows;
    c.R1 = c.R0 + a.band_rows < c.H ? c.R0 + a.band_rows : c.H;
    c.K0 = c.R0 / 2; c.K1 = c.R1 / 2;
    c.M = (c.R1 - c.R0) + 7 + BW_DF;
    c.Wp = 4 * GL;
    c.HW = (size_t)c.H * c.W;
    if (a.n_strips > 1) {
        c.v0 = strip * BW_STRIP;
        c.v1 = c.v0 + BW_STRIP < c.W ? c.v0 + BW_STRIP : c.W;
        c.x0 = strip ? c.v0 - 8 : 0;
    } else {
        c.x0 = 0; c.v0 = 0; c.v1 = c.W;
    }
    return c;
}

// warp-reduce one value and add lane 0's total to a global accumulator
__device__ __forceinline__ void warp_atomic(float v, float* dst) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(dst, v);
}

template <int MODE, bool XW, bool ISL, bool FINE>
__device__ __forceinline__ void bw_walk(const StreamBwdArgs& a, float* smem, const int wk, const int nwk, const int lane, const int GL) {
    using SM = BwSmem<MODE>;
    constexpr bool HAS_L = SM::HAS_L, THR = SM::THR, RAW = ISL || THR;
    constexpr int NSRC = SM::NSRC, NOP = SM::NOP, NK = SM::NK, ZR = BW_ZR, PD = BW_PD;
    constexpr int KIND = ISL ? 1 : 0;
    constexpr int PL0 = ISL ? SM::NPT : 0, NMY = ISL ? 4 : SM::NPT;          // this kind's planes in the weight rings
    constexpr bool FIN = FINE && !ISL;                                       // the finisher role
    constexpr bool XWF = XW && FINE;
    const BwCta ct = bw_cta(a, GL);
    const int H = ct.H, W = ct.W, G = ct.G, nch = ct.nch, g = ct.g, R0 = ct.R0, R1 = ct.R1, M = ct.M;
    const bool live = wk < nch;
    const int wkc = live ? wk : 0;
    const int c = g * ct.F + ct.f0 + wkc;
    const size_t off = ((size_t)ct.b * G * ct.F + c) * ct.HW;
    const size_t plane = (size_t)ct.b * G + g;

    SM lay; lay.Wp = ct.Wp; lay.nch = nch;
    const int Wp = ct.Wp, Wpc = Wp / 2;
    const int LH = FINE ? H : H / 2, LW = FINE ? W : W / 2, LHW = LH * LW;
    const int wpitch = FINE ? Wp : Wpc;
    const float* zring = smem + lay.zring() + (size_t)wkc * ZR * Wp;
    const float* s0ring = smem + lay.sring() + (size_t)wkc * ZR * Wp;
    const float* s1ring = s0ring + (size_t)nch * ZR * Wp;
    const float* opring = smem + lay.opring() + (size_t)wkc * NOP * BW_OPR * Wp;
    float* xring = smem + lay.xring() + (size_t)wkc * 2 * NK * 2 * Wp;
    float* cring = smem + lay.cring() + (size_t)wkc * 4 * NK * 2 * Wpc;
    float* mbox = smem + lay.mbox();
    const smem_addr_t sbase = smem_addr(smem);

    // ---- per-graph scalars and this walker's module
    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const bool has_skip = a.p.skip != nullptr;
    const float s0 = has_skip ? a.p.skip[0] : 0.f, s1 = has_skip ? a.p.skip[1] : 1.f;
    const float c23 = al2 * s1;
    float ca, cb = 0.f;      // g = ca * src0 + cb * src1
    if (MODE == BW_X3) ca = -c23;
    else if (MODE == BW_X2A) { ca = -be2 * c23; cb = -al1; }
    else if (MODE == BW_X2B) { ca = c23 + be2 * c23; cb = al1; }
    else if (MODE == BW_X1) ca = -al0;
    else ca = 1.f;
    const glrgtv_opparams& op = ISL ? (FINE ? a.p.glr0 : a.p.glr1) : (FINE ? a.p.gtv0 : a.p.gtv1);
    const StatsTaps kK = glr_load_taps(op.stats, c);
    const float aK = expf(ISL ? (FINE ? a.p.mu0[g] : a.p.mu1[g]) : (FINE ? a.p.ro0[g] : a.p.ro1[g]));
    const float Gam = THR ? expf(FINE ? a.p.gamma0[g] : a.p.gamma1[g]) : 0.f;

    LaneCtx lc;
    const int lcol = 4 * lane;                                  // column inside the walker's window (shared-memory rings)
    lc.col0 = (FINE ? ct.x0 : ct.x0 / 2) + lcol;                // column in the plane of this resolution
    // columns whose results count (a strip's halo lanes compute throw-away values)
    const bool valid = FINE ? (lc.col0 >= ct.v0 && lc.col0 < ct.v1) : (lc.col0 >= ct.v0 / 2 && lc.col0 < ct.v1 / 2);
    lc.width = FINE ? (GL < 32 ? GL : 32) : (GL / 2 < 32 ? GL / 2 : 32);
    lc.active = live && lc.col0 < LW;
    lc.first = lc.col0 == 0;
    lc.last = lc.col0 + 4 >= LW;
    lc.seam_l = XWF && live && lane == 32 && lc.col0 < LW;
    lc.seam_r = XWF && live && lane == 31 && lc.col0 + 4 < LW;
    lc.mb_rd = lc.mb_wr = mbox;

    // ---- cp.async loader: this thread's share, BW_PD block steps ahead
    //      weights: planes wk, wk+nwk, ... of this walker's kind and level; coarse T lanes: z and the upstream sources;
    //      fine T lanes: the finisher's operand rows
    const float* wsrc[BW_MAXJ];
    smem_addr_t wdst[BW_MAXJ];
    bool wlead[BW_MAXJ];
    const float* wring0 = smem + (FINE ? lay.w0ring() : lay.w1ring()) + (size_t)PL0 * BW_WR * wpitch + lcol;   // my kind's plane 0, my quad
    {
        const float* base = ISL ? (FINE ? a.wL0 : a.wL1) + plane * 4 * LHW
                                : THR ? (FINE ? a.wT0 : a.wT1) + plane * 4 * LHW : (FINE ? a.cT0 : a.cT1) + plane * 2 * LHW;
#pragma unroll
        for (int j = 0; j < BW_MAXJ; ++j) {
            const int e = wk + j * nwk;
            wsrc[j] = nullptr; wdst[j] = sbase; wlead[j] = false;
            if (lc.col0 < LW && e < NMY) {
                wsrc[j] = base + (size_t)e * LHW + lc.col0;
                wdst[j] = smem_advance(sbase, (int)(FINE ? lay.w0ring() : lay.w1ring()) + (PL0 + e) * BW_WR * wpitch + lcol);
                wlead[j] = RAW && e == 0;
            }
        }
    }
    const float* opp[3] = {nullptr, nullptr, nullptr};
    if (FIN && lc.active) {
        if (NOP > 0) opp[0] = a.op0 + off + lc.col0;
        if (NOP > 1) opp[1] = a.op1 + off + lc.col0;
        if (NOP > 2 && has_skip) opp[2] = a.op2 + off + lc.col0;
    }
    const smem_addr_t opdst = smem_advance(sbase, (int)lay.opring() + wkc * NOP * BW_OPR * Wp + lcol);
    const smem_addr_t zdst = smem_advance(sbase, (int)lay.zring() + wkc * ZR * Wp + 2 * lcol);
    const smem_addr_t s0dst = smem_advance(sbase, (int)lay.sring() + wkc * ZR * Wp + 2 * lcol);
    const smem_addr_t s1dst = smem_advance(s0dst, nch * ZR * Wp);
    int zis = 0;
    const int r0 = FINE ? R0 : ct.K0;
    auto issue = [&](int mt) {
        if (mt >= M) return;
        if (FINE ? (mt < BW_DF) : (mt & 1)) return;
        const int t = r0 - 3 + (FINE ? mt - BW_DF : (mt >> 1));
        if (!FINE && !ISL) {
            if (lc.active) {
                const int rho = 2 * t;
                const size_t go = off + (size_t)rho * W + 2 * lc.col0;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (rho + k < 0 || rho + k >= H) continue;
                    const int so = (zis + k) * Wp;
                    cp_async16_s(smem_advance(zdst, so), a.z + go + k * W); cp_async16_s(smem_advance(zdst, so + 4), a.z + go + k * W + 4);
                    cp_async16_s(smem_advance(s0dst, so), a.src0 + go + k * W); cp_async16_s(smem_advance(s0dst, so + 4), a.src0 + go + k * W + 4);
                    if (NSRC == 2) { cp_async16_s(smem_advance(s1dst, so), a.src1 + go + k * W); cp_async16_s(smem_advance(s1dst, so + 4), a.src1 + go + k * W + 4); }
                }
            }
            zis = zis + 2 == ZR ? 0 : zis + 2;
        }
        if (FIN && NOP > 0) {
            const int re = t - 4;                          // the finisher's row
            if (re >= R0 && re < R1) {
                const int so = (re & (BW_OPR - 1)) * Wp, go = re * W;
#pragma unroll
                for (int k = 0; k < NOP; ++k)
                    if (opp[k] != nullptr) cp_async16_s(smem_advance(opdst, k * BW_OPR * Wp + so), opp[k] + go);
            }
        }
        const int rw = t - 2;
        const bool ok0 = rw >= 0 && rw < LH, ok1 = rw + 1 >= 0 && rw + 1 < LH;
        const int so0 = (rw & (BW_WR - 1)) * wpitch, so1 = ((rw + 1) & (BW_WR - 1)) * wpitch;
        const int go0 = rw * LW, go1 = go0 + LW;
#pragma unroll
        for (int j = 0; j < BW_MAXJ; ++j) {
            if (wsrc[j] == nullptr) continue;
            if (wlead[j]) { if (ok1) cp_async16_s(smem_advance(wdst[j], so1), wsrc[j] + go1); }
            else if (ok0) cp_async16_s(smem_advance(wdst[j], so0), wsrc[j] + go0);
        }
    };
    auto wrow = [&](int e, int row) { return wring0 + (e * BW_WR + (row & (BW_WR - 1))) * wpitch; };

    Row z[3], gq[3], s[3], h[3], o[3], gs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) z[k] = gq[k] = s[k] = h[k] = o[k] = gs[k] = row_zero();
    Row z3 = row_zero(), g3 = row_zero(), cDp = row_zero(), wU_next = row_zero(), wD_prev = row_zero();
    float stt[5] = {0.f, 0.f, 0.f, 0.f, 0.f};     // tap sums (c, R, D, U, L) of this walker's stats kernel
    float sumK = 0.f;                             // mu / ro
    float fs[4] = {0.f, 0.f, 0.f, 0.f};           // finisher: alpha_k, beta2, skip0, skip1
    int zs = FINE ? 3 : 0;                        // ring slot of fine row t (fine) / of fine row 2t (coarse)

#pragma unroll
    for (int k = 0; k < PD; ++k) { issue(k); cp_async_commit(); }

    // One step body per role: with four roles in the kernel, phase-unrolled windows (as in the forward kernels) blow the
    // instruction cache (measured: "no instruction" was the top stall); the windows are rotated with register moves instead.
#pragma unroll 1
    for (int m = 0; m < M; ++m) {
        {
            cp_async_wait_pending<PD - 1>();
            __syncthreads();
            issue(m + PD);
            cp_async_commit();
            if (FINE ? (m < BW_DF) : (m & 1)) continue;
            constexpr int N = 2, C = 1, U = 0;
            const int t = r0 - 3 + (FINE ? m - BW_DF : (m >> 1));
            if (XWF) {
                lc.mb_rd = mbox + (((m + 1) & 1) * nch * NK + wkc * NK + KIND) * BP_COUNT * 2;
                lc.mb_wr = mbox + ((m & 1) * nch * NK + wkc * NK + KIND) * BP_COUNT * 2;
            }
            auto zslot = [&](int back) { int v = zs - back; return v < 0 ? v + ZR : v; };     // ring slot of fine row t - back
            // ---- rows t of z (clamp-extended at production) and of the upstream g (zero outside)
            if (t >= 0 && t < LH) {
                if (FINE) {
                    const int so = zs * Wp + lcol;
                    z[N] = row_ld(zring + so);
                    const Row q0 = row_ld(s0ring + so);
                    Row q1 = row_zero();
                    if (NSRC == 2) q1 = row_ld(s1ring + so);
#pragma unroll
                    for (int j = 0; j < 4; ++j) gq[N].v[j] = ca * q0.v[j] + cb * q1.v[j];
                } else {
                    const int so = zs * Wp + 2 * lcol;
                    auto pool = [&](const float* ring, Row& dst) {
                        const float* p0 = ring + so;
                        const Row a0 = row_ld(p0), a1 = row_ld(p0 + 4), b0 = row_ld(p0 + Wp), b1 = row_ld(p0 + Wp + 4);
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
                    for (int j = 0; j < 4; ++j) gq[N].v[j] = ca * q0.v[j] + cb * q1.v[j];
                }
                if (t == 0) z[C] = z[N];
            } else {
                z[N] = t < 0 ? row_zero() : z[C];
                gq[N] = row_zero();
            }
            mb_post<XWF>(gq[N], lc, BP_G);
            // scalars left / right of a z row that is still in the ring (fine) or in registers (coarse)
            auto z_lr = [&](const Row& zc, int back, float& l, float& r) {
                if (FINE) {
                    const float* p = zring + zslot(back) * Wp + lcol;
                    // (the first / last lane of a column strip inside the image owns halo columns whose results are thrown away: it
                    // takes its own value instead of reading the ring element next door - for row slot 0 of channel 0 that would be
                    // the word BEFORE the shared-memory block, tools/emu_asan.sh)
                    l = (lc.first || lcol == 0) ? zc.v[0] : p[-1];
                    r = (lc.last || lcol + 4 >= Wp) ? zc.v[3] : p[4];
                } else {
                    nb_lr<false, false>(zc, l, r, lc, 0);
                }
            };
            // ---- first-level stencils at row t-1: s = S z (clamp-extended), h = a S0 g (T: clamp-extended, L: zero-extended)
            {
                const int r = t - 1;
                if (r >= 0 && r < LH) {
                    float l, rr;
                    z_lr(z[C], 1, l, rr);
                    s[N] = w_S(kK, z[C], z[U], z[N], l, rr);
                    nb_lr<true, XWF>(gq[C], l, rr, lc, BP_G);
                    h[N] = w_S(kK, gq[C], gq[U], gq[N], l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[N].v[j] *= aK;
                    if (r == 0) { s[C] = s[N]; if (!ISL) h[C] = h[N]; }
                } else if (r >= LH) {
                    s[N] = s[C];
                    h[N] = ISL ? row_zero() : h[C];
                } else {
                    h[N] = row_zero();
                }
                mb_post<XWF>(s[N], lc, BP_S);
                mb_post<XWF>(h[N], lc, BP_H);
            }
            // ---- cores at row t-2: o = core(s), gs = core'(h); stats gradient of S: sum gs * z[5-point]
            {
                const int r = t - 2;
                if (r >= 0 && r < LH) {
                    float sl, sr, hl, hr;
                    nb_lr<false, XWF>(s[C], sl, sr, lc, BP_S);
                    if (ISL) {
                        nb_lr<true, XWF>(h[C], hl, hr, lc, BP_H);
                        RawW rw;
                        ring_raw_w(wring0, wpitch, r, LH, lc.first, lc.last, wU_next, wD_prev, rw);
                        o[N] = w_L(s[C], s[U], s[N], sl, sr, rw.own);
                        gs[N] = w_L_adj(h[C], h[U], h[N], hl, hr, rw, r == 0, r == LH - 1, lc.first, lc.last);
                    } else if (THR) {
                        nb_lr<false, XWF>(h[C], hl, hr, lc, BP_H);
                        RawW rw;
                        ring_raw_w(wring0, wpitch, r, LH, lc.first, lc.last, wU_next, wD_prev, rw);
                        o[N] = w_core_thr(s[C], s[U], s[N], sl, sr, rw, Gam);
                        gs[N] = w_core_thr_adj(h[C], h[U], h[N], hl, hr, s[C], s[U], s[N], sl, sr, rw, Gam);
                    } else {
                        nb_lr<false, XWF>(h[C], hl, hr, lc, BP_H);
                        const float* crow = wrow(0, r);
                        const Row cr = row_ld(crow), cd = row_ld(wrow(1, r));
                        const float cr_left = lc.first ? 0.f : crow[-1];
                        o[N] = w_core_lin(s[C], s[U], s[N], sl, sr, cr, cr_left, cd, cDp);
                        gs[N] = w_core_lin(h[C], h[U], h[N], hl, hr, cr, cr_left, cd, cDp);     // the linear core is self-adjoint
                        cDp = cd;
                    }
                    float zl, zr;
                    z_lr(z[U], 2, zl, zr);           // z rows t-3 (z3), t-2 (z[U]), t-1 (z[C]); (shuffles: whole warp)
                    if (lc.active && valid && (FINE ? (r >= R0 && r < R1) : (r >= ct.K0 && r < ct.K1))) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float gv = gs[N].v[j];
                            stt[0] += gv * z[U].v[j];
                            stt[1] += gv * rowR(z[U], zr, j);
                            stt[2] += gv * z[C].v[j];
                            stt[3] += gv * z3.v[j];
                            stt[4] += gv * rowL(z[U], zl, j);
                        }
                    }
                } else {
                    o[N] = row_zero(); gs[N] = row_zero(); cDp = row_zero();
                    if (RAW) {
                        wD_prev = row_zero();
                        wU_next = (r + 1 >= 0 && r + 1 < LH) ? row_ld(wrow(0, r + 1)) : row_zero();
                    }
                }
                mb_post<XWF>(o[N], lc, BP_O);
                mb_post<XWF>(gs[N], lc, BP_GS);
            }
            // ---- St / S' at row t-3: fwd = a St o, V = S' gs; stats gradient of St: sum (a g) * o[5-point]; mu / ro
            {
                const int r = t - 3;
                if (FINE ? (r >= R0 && r < R1) : (r >= 0 && r < LH)) {
                    float l, rr;
                    nb_lr<true, XWF>(o[C], l, rr, lc, BP_O);
                    Row fwd = w_St(kK, o[C], o[U], o[N], l, rr);
                    const bool cnt = lc.active && valid && (FINE || (r >= ct.K0 && r < ct.K1));
                    if (cnt) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float gu = aK * g3.v[j];
                            stt[0] += gu * o[C].v[j];
                            stt[1] += gu * rowL(o[C], l, j);
                            stt[2] += gu * o[U].v[j];
                            stt[3] += gu * o[N].v[j];
                            stt[4] += gu * rowR(o[C], rr, j);
                            sumK += gu * fwd.v[j];
                        }
                    }
                    nb_lr<true, XWF>(gs[C], l, rr, lc, BP_GS);
                    Row V = w_St(kK, gs[C], gs[U], gs[N], l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float self = 0.f;
                        if (lc.last && j == 3) self += kK.kr;
                        if (r == LH - 1) self += kK.kd;
                        if (r == 0) self += kK.ku;
                        if (lc.first && j == 0) self += kK.kl;
                        V.v[j] += self * gs[C].v[j];
                        fwd.v[j] *= aK;
                    }
                    if (lc.active) {
                        if (FINE) {
                            float* slot = xring + (((m & 1) * NK + KIND) * 2) * Wp + lcol;
                            st4(slot, V.v);
                            st4(slot + Wp, fwd.v);
                        } else {
                            float* slot = cring + (((r & 3) * NK + KIND) * 2) * Wpc + lcol;
#pragma unroll
                            for (int j = 0; j < 4; ++j) { V.v[j] *= 0.25f; fwd.v[j] *= 0.25f; }
                            st4(slot, V.v);
                            st4(slot + Wpc, fwd.v);
                        }
                    }
                }
            }
            // ---- finisher: row t-4, everything its walkers posted one step ago
            if (FIN) {
                const int r = t - 4;
                if (r >= R0 && r < R1 && lc.active && valid) {
                    const float* xs = xring + (((m + 1) & 1) * NK * 2) * Wp + lcol;
                    const float* cs = cring + (((r >> 1) & 3) * NK * 2) * Wpc + (lcol >> 1);
                    Row V = row_ld(xs), Fw = row_ld(xs + Wp);
                    GLR_CHECK_ALIGN(cs, 8);
                    float2 cV = *reinterpret_cast<const float2*>(cs), cF = *reinterpret_cast<const float2*>(cs + Wpc);
                    if (HAS_L) {
                        const Row VL = row_ld(xs + 2 * Wp), FL = row_ld(xs + 3 * Wp);
                        const float2 cVL = *reinterpret_cast<const float2*>(cs + 2 * Wpc), cFL = *reinterpret_cast<const float2*>(cs + 3 * Wpc);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { V.v[j] += VL.v[j]; Fw.v[j] += FL.v[j]; }
                        cV.x += cVL.x; cV.y += cVL.y; cF.x += cFL.x; cF.y += cFL.y;
                    }
                    const int so = zslot(4) * Wp + lcol;
                    const Row zq = row_ld(zring + so), q0 = row_ld(s0ring + so);
                    Row q1 = row_zero();
                    if (NSRC == 2) q1 = row_ld(s1ring + so);
                    const float* ops = opring + (r & (BW_OPR - 1)) * Wp + lcol;
                    Row p0 = row_zero(), p1 = row_zero(), p2 = row_zero(), outv;
                    if (NOP > 0) p0 = row_ld(ops);
                    if (NOP > 1) p1 = row_ld(ops + BW_OPR * Wp);
                    if (NOP > 2 && has_skip) p2 = row_ld(ops + 2 * BW_OPR * Wp);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float gj = ca * q0.v[j] + cb * q1.v[j];
                        const float Vj = (HAS_L ? gj : 0.f) + V.v[j] + (j < 2 ? cV.x : cV.y);
                        const float Az = zq.v[j] + Fw.v[j] + (j < 2 ? cF.x : cF.y);
                        if (MODE == BW_X3) {            // q0 gout, p0 r1, p1 bB, p2 x
                            const float u2 = (p1.v[j] - Az) + be2 * p0.v[j], x3 = zq.v[j] + al2 * u2, gg = s1 * q0.v[j];
                            outv.v[j] = gg + Vj;
                            fs[0] += gg * u2;
                            fs[1] += al2 * gg * p0.v[j];
                            if (has_skip) { fs[2] += q0.v[j] * p2.v[j]; fs[3] += q0.v[j] * x3; }
                        } else if (MODE == BW_X2A) {    // q1 gx2, p0 r1
                            outv.v[j] = q1.v[j] + Vj;
                            fs[0] += q1.v[j] * p0.v[j];
                        } else if (MODE == BW_X2B) {    // p0: the A part already in gx1
                            outv.v[j] = p0.v[j] + Vj;
                        } else if (MODE == BW_X1) {     // q0 gx1
                            outv.v[j] = (1.f + al0) * q0.v[j] + Vj;
                            fs[0] += q0.v[j] * (zq.v[j] - Az);
                        } else {                        // BA: q0 gbA, p0 gout, p1 gx2
                            const float gr2 = c23 * p0.v[j];
                            outv.v[j] = q0.v[j] + Vj + (gr2 + be2 * gr2 + al1 * p1.v[j]) + s0 * p0.v[j];
                        }
                    }
                    st4(a.out + off + (size_t)r * W + lc.col0, outv.v);
                }
            }
            if (FINE) zs = zs + 1 == ZR ? 0 : zs + 1;
            else zs = zs + 2 == ZR ? 0 : zs + 2;
            // ---- rotate the windows (rows t-2 of z and g stay one more step: the stats sums read them as row t-3)
            z3 = z[U]; g3 = gq[U];
            z[U] = z[C]; z[C] = z[N]; gq[U] = gq[C]; gq[C] = gq[N]; s[U] = s[C]; s[C] = s[N];
            h[U] = h[C]; h[C] = h[N]; o[U] = o[C]; o[C] = o[N]; gs[U] = gs[C]; gs[C] = gs[N];
        }
    }
    cp_async_wait_all();
    // ---- parameter gradients of this walker: tap sums -> p01, p02a, p02b, p03 (linear), then one atomic per warp and value
    {
        float* dst = ISL ? (FINE ? a.gr.glr0_stats : a.gr.glr1_stats) : (FINE ? a.gr.gtv0_stats : a.gr.gtv1_stats);
        const int C = G * ct.F;
        const float v4[4] = {stt[0], stt[1] - stt[0], stt[2] - stt[0], 4.f * stt[0] - stt[1] - stt[2] - stt[3] - stt[4]};
        // lanes of one warp may belong to different channels (walkers narrower than a warp): reduce per walker
        const int gl = FINE ? GL : GL / 2;
        if (gl >= 32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) warp_atomic(live ? v4[k] : 0.f, dst + k * C + c);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v = live ? v4[k] : 0.f;
                for (int o2 = gl / 2; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
                if (lane == 0 && live && v != 0.f) atomicAdd(dst + k * C + c, v);
            }
        }
        float* sdst = ISL ? (FINE ? a.gr.mu0 : a.gr.mu1) : (FINE ? a.gr.ro0 : a.gr.ro1);
        warp_atomic(live ? sumK : 0.f, sdst + g);
        if (FIN) {
            const int kA = MODE == BW_X3 ? 2 : MODE == BW_X2A ? 1 : 0;
            if (MODE == BW_X3 || MODE == BW_X2A || MODE == BW_X1) warp_atomic(fs[0], a.gr.alpha + kA * G + g);
            if (MODE == BW_X3) {
                warp_atomic(fs[1], a.gr.beta + 2 * G + g);
                if (has_skip && a.gr.skip) { warp_atomic(fs[2], a.gr.skip); warp_atomic(fs[3], a.gr.skip + 1); }
            }
        }
    }
}

// threads per CTA / resident CTAs the register budget is sized for.  The thresholded stage X2B stages two source tensors
// and four raw weight planes per level: two of its CTAs do not fit one SM's shared memory, so it runs ONE larger CTA.
template <int MODE> struct BwLaunch { static constexpr int MAXT = MODE == BW_X2B ? 288 : BW_MAXT, MINB = MODE == BW_X2B ? 1 : 2; };

template <int MODE, bool XW>
__global__ void __launch_bounds__(BwLaunch<MODE>::MAXT, BwLaunch<MODE>::MINB) k_stream_bwd(StreamBwdArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool HAS_L = BwSmem<MODE>::HAS_L;
    const int W = a.s.W;
    const int GL = XW ? 64 : (W <= 32 ? 8 : W <= 64 ? 16 : 32), GLc = GL / 2;
    const int NF = (a.nch * GL + 31) & ~31, NC = (a.nch * GLc + 31) & ~31;     // threads of one fine / coarse role
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    {
        BwSmem<MODE> lay; lay.Wp = 4 * GL; lay.nch = a.nch;
        const int n4 = (int)(lay.total() / 4);
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += NT) st4(smem + 4 * i, z4);
        __syncthreads();
    }
    // roles in thread order: fine T | fine L | coarse T | coarse L   (L only where the stage has a GLR part).  Role boundaries are
    // whole warps; the warp index is broadcast from lane 0 so that the compiler KNOWS the role branch is warp-uniform (otherwise
    // every shuffle inside a role is bracketed by convergence barriers)
    const int bFL = NF, bCT = HAS_L ? 2 * NF : NF, bCL = bCT + NC;
    const int wtid = 32 * (int)__shfl_sync(0xffffffffu, (float)(tid >> 5), 0);
    if (wtid < bFL) bw_walk<MODE, XW, false, true>(a, smem, tid / GL, NF / GL, tid % GL, GL);
    else if (HAS_L && wtid < bCT) bw_walk<MODE, XW, true, true>(a, smem, (tid - bFL) / GL, NF / GL, (tid - bFL) % GL, GL);
    else if (!HAS_L || wtid < bCL) bw_walk<MODE, XW, false, false>(a, smem, (tid - bCT) / GLc, NC / GLc, (tid - bCT) % GLc, GL);
    else bw_walk<MODE, XW, true, false>(a, smem, (tid - bCL) / GLc, NC / GLc, (tid - bCL) % GLc, GL);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
extern unsigned long long g_glr_stream_launches;
struct BwPlan {
    int GL, nch, threads, band_rows, n_bands, n_strips;
};
template <int MODE>
static BwPlan bw_plan(const glrgtv_shape& s) {
    constexpr bool HAS_L = BwSmem<MODE>::HAS_L;
    BwPlan p;
    p.GL = s.W > 128 ? 64 : s.W > 64 ? 32 : s.W > 32 ? 16 : 8;
    p.n_strips = s.W > 256 ? (s.W + BW_STRIP - 1) / BW_STRIP : 1;
    auto threads = [&](int n) {
        const int nf = (n * p.GL + 31) & ~31, nc = (n * p.GL / 2 + 31) & ~31;
        return (HAS_L ? 2 : 1) * (nf + nc);
    };
    // channels per CTA: as many as the thread budget and shared memory allow (measured on B200: sharing the staged
    // weight rows between more channel walkers beats the extra resident CTAs of a smaller choice)
    p.nch = 1;
    for (int n = 1; n <= s.F; ++n) {
        if (s.F % n || threads(n) > BwLaunch<MODE>::MAXT) continue;
        BwSmem<MODE> lay; lay.Wp = 4 * p.GL; lay.nch = n;
        if (lay.total() * sizeof(float) + 1024 > 227 * 1024) continue;
        p.nch = n;
    }
    p.threads = threads(p.nch);
    const long ctas = (long)s.B * s.G * (s.F / p.nch) * p.n_strips;
    int bands = 1;
    while (ctas * bands < 296 && s.H / (bands * 2) >= 32) bands *= 2;
    p.band_rows = ((s.H + bands - 1) / bands + 1) & ~1;
    p.n_bands = (s.H + p.band_rows - 1) / p.band_rows;
    return p;
}

template <int MODE, bool XW>
static int launch_bw_kernel(const StreamBwdArgs& a, const BwPlan& p, long blocks, void* stream) {
    BwSmem<MODE> lay; lay.Wp = 4 * p.GL; lay.nch = p.nch;
    const size_t smem = lay.total() * sizeof(float);
    if (smem > 227 * 1024) return GLRGTV_ERR_UNSUPPORTED;
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_stream_bwd<MODE, XW>, smem, optin)) return rc_;
#endif
    ++g_glr_stream_launches;
    GLR_LAUNCH_FIBERS((k_stream_bwd<MODE, XW>), dim3((unsigned)blocks), p.threads, smem, stream, a);
    return GLRGTV_OK;
}
template <int MODE>
int glr_stream_bwd_stage(StreamBwdArgs a, int slot, void* stream) {
    const glrgtv_shape& s = a.s;
    const BwPlan p = bw_plan<MODE>(s);
    a.nch = p.nch; a.band_rows = p.band_rows; a.n_bands = p.n_bands; a.n_strips = p.n_strips;
    const long blocks = (long)s.B * s.G * (s.F / p.nch) * p.n_bands * p.n_strips;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    GLR_PROF_BEGIN(slot, stream);
    const int rc = p.GL == 64 ? launch_bw_kernel<MODE, true>(a, p, blocks, stream) : launch_bw_kernel<MODE, false>(a, p, blocks, stream);
    GLR_PROF_END(slot, stream);
    return rc ? rc : GLR_CHECK_LAUNCH();
}
template int glr_stream_bwd_stage<BW_X3>(StreamBwdArgs, int, void*);
template int glr_stream_bwd_stage<BW_X2A>(StreamBwdArgs, int, void*);
template int glr_stream_bwd_stage<BW_X2B>(StreamBwdArgs, int, void*);
template int glr_stream_bwd_stage<BW_X1>(StreamBwdArgs, int, void*);
template int glr_stream_bwd_stage<BW_BA>(StreamBwdArgs, int, void*);
