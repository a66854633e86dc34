// fw2.cuh - second-generation streaming FORWARD stages of LocalLowpassFilteringBlock / MixtureGTVGLR (V1X0:707-811, 985-988):
// the pair-lane walker of bw2.cuh without the adjoint chain.
//
// Same execution model as the backward (bw2.cuh): a lane owns a pair of adjacent pixels (FFMA2 arithmetic), one walker per
// channel does both operator families of ONE resolution, the half-resolution branch is its own launch (COARSE) run first,
// whose result vc = a_T1 Ft1(P z) + a_L1 Fl1(P z) the full-resolution launch adds as 0.25 P^T vc in its epilogue; rows of z and
// of every weight plane are staged two steps ahead by cp.async into shared-memory rings (clamped / zero-filled outside the
// image), St runs in scatter form.  Per step t: stencils S at row t-1, cores at row t-2, St fed with row t-3, epilogue at t-4.
//
//   stage   input   result                                                                          V1X0
//   BA      y       bA = y + R_lin y                         R = a_T St_T K_T S_T (+ half res.)     738-749
//   X1      bA      x1 = bA + a0 (bA - A bA)                 A = I + R_lin + a_L St_L L S_L          751-753
//   X2      x1      bB = y + R_thr x1; r1 = bB - A x1; x2 = x1 + a1 r1                               757-786
//   X3      x2      u2 = bB - A x2 + b2 r1; out = s0 x + s1 (x2 + a2 u2)                             788-790, 985-988
//
// Planes wider than the walker (4K inference) are cut into column strips: a strip's window carries 8 halo columns per side
// (the reach of one stage is 3 pixels; 8 keeps every 16-byte copy of the half-resolution result aligned), its edge lanes
// compute throw-away values, and stores are masked to the strip's own columns.  A launch may be restricted to a row range
// (glrgtv_block_fwd_stage: spatially sharded inference).
#pragma once
#include "bw2.cuh"

enum { FW_BA = 0, FW_X1 = 1, FW_X2 = 2, FW_X3 = 3 };
enum { F2_ST = 0, F2_SL = 1, F2_OT = 2, F2_OL = 3, F2_OH = 4, F2_NFLD = B2_NFLD };      // seam mailbox fields (OH: thresholded core); slot size as in bw2.cuh
#define F2_HALO 8

struct F2Args {
    glrgtv_shape s;             // FULL-resolution geometry of the block
    glrgtv_block_params p;
    const float* z;             // y | bA | x1 | x2                         [B,C,H,W]
    const float* op0;           // fine: X2 y | X3 bB
    const float* op1;           // fine: X3 r1
    const float* op2;           // fine: X3 x (skip path; may be the same tensor as y)
    const float* vc_in;         // fine: half-resolution results [nvc][B,C,H/2,W/2] (X2: linear part, thresholded part)
    float* vc_out;              // coarse
    const float* cT;            // this level: symmetric GTV coefficients [B,G,2,LH,LW]
    const float* wT;            // this level: raw GTV weights (X2)       [B,G,4,LH,LW]
    const float* wL;            // this level: GLR weights                [B,G,4,LH,LW]
    float* out0; float* out1; float* out2;      // BA: bA | X1: x1 | X2: x2, bB, r1 | X3: out
    int lg, nch, n_parts;       // log2(floats per shared-memory row), channels per CTA, CTAs per graph
    int row0, row1;             // level rows this launch produces
    int band_rows, n_bands;
    int n_strips, strip_w;      // column strips: valid level columns per strip (the whole width when n_strips == 1)
};

template <int MODE, bool COARSE>
struct F2Smem {
    static constexpr bool HAS_L = MODE != FW_BA, THR = MODE == FW_X2;
    static constexpr int NPT = 2 + (THR ? 4 : 0), NPLW = NPT + (HAS_L ? 4 : 0);       // cT(2) [raw wT(4)] [wL(4)]
    static constexpr int NOP = COARSE ? 0 : MODE == FW_X2 ? 1 : MODE == FW_X3 ? 3 : 0;
    static constexpr int NVC = THR ? 2 : 1;
    int NCH, L;
    __host__ __device__ int rw() const { return 2 * L; }
    __host__ __device__ size_t zsig() const { return 0; }                                                    // [NCH][ZR][RW]
    __host__ __device__ size_t zstage() const { return zsig() + (size_t)NCH * B2_ZR * rw(); }                // COARSE [NCH][ZRC][2][2 RW]
    __host__ __device__ size_t opring() const { return zstage() + (COARSE ? (size_t)NCH * B2_ZRC * 4 * rw() : 0); }  // [NOP][NCH][4][RW]
    __host__ __device__ size_t vcring() const { return opring() + (size_t)NOP * NCH * B2_OPR * rw(); }       // fine [NVC][NCH][4][RW/2]
    __host__ __device__ size_t wring() const { return vcring() + (COARSE ? 0 : (size_t)NVC * NCH * B2_OPR * (rw() / 2)); }   // [NPLW][WR][RW]
    __host__ __device__ size_t mbox() const { return wring() + (size_t)NPLW * B2_WR * rw(); }                // [2][NCH][warps per row][NFLD][2]
    __host__ __device__ size_t total() const { return (mbox() + (size_t)2 * NCH * ((L + 31) / 32) * F2_NFLD * 2 + 8 + 3) & ~(size_t)3; }
    __host__ __device__ size_t bytes() const { return total() * sizeof(float); }
};

template <int MODE, bool COARSE, int LGT, int NCHT, bool XWG>
__global__ void __launch_bounds__(B2_MAXT, MODE == FW_X2 ? 1 : 2) k_fw2(F2Args a) {      // light stages: two CTAs per SM
    GLR_SMEM_DECL(smem);
    using SM = F2Smem<MODE, COARSE>;
    constexpr bool HAS_L = SM::HAS_L, THR = SM::THR;
    constexpr bool XW = LGT ? (LGT >= 7) : XWG;
    constexpr int NPT = SM::NPT, NPLW = SM::NPLW, NOP = SM::NOP, NVC = SM::NVC, PD = B2_PD;
    constexpr int PL_RAW = 2, PL_L = NPT;                    // first plane of the raw GTV set / of the GLR set in the weight rings
    const int W = a.s.W, F = a.s.F, G = a.s.G;
    const int LH = COARSE ? a.s.H / 2 : a.s.H, LW = COARSE ? W / 2 : W;
    const int LG = LGT ? LGT : a.lg, RW = 1 << LG, L = RW >> 1, NCH = NCHT ? NCHT : a.nch;
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ch = tid >> (LG - 1), lr = tid & (L - 1);
    const bool live = ch < NCH;
    const int chc = live ? ch : 0;
    int bid = (int)blockIdx.x;
    const int strip = bid % a.n_strips; bid /= a.n_strips;
    const int band = bid % a.n_bands; bid /= a.n_bands;
    const int part = bid % a.n_parts; bid /= a.n_parts;
    const int g = bid % G, b = bid / G;
    const int R0 = a.row0 + band * a.band_rows, R1 = R0 + a.band_rows < a.row1 ? R0 + a.band_rows : a.row1;
    const int M = (R1 - R0) + 7;
    const int c = g * F + part * NCH + chc;
    const size_t pl_f = (size_t)b * G * F + c;
    const size_t plane = (size_t)b * G + g;
    // column window of this strip: valid level columns [v0, v1), staged from x0 (8 halo columns inside the image)
    const int v0 = strip * a.strip_w, v1 = v0 + a.strip_w < LW ? v0 + a.strip_w : LW;
    const int x0 = v0 >= F2_HALO ? v0 - F2_HALO : 0;
    const int col0 = 2 * lr, gcol = x0 + col0;               // column inside the window / in the plane
    const int NWR = (L + 31) >> 5;

    SM lay; lay.NCH = NCH; lay.L = L;
    float* const zsig = smem + lay.zsig() + ((size_t)chc * B2_ZR << LG) + col0;
    float* const wring = smem + lay.wring() + col0;
    const int M8 = (B2_ZR << LG) - 1, M4 = (B2_WR << LG) - 1;

    B2Lane lc;
    lc.width = L < 32 ? L : 32;
    const bool active = live && gcol < LW;
    const bool valid = active && gcol >= v0 && gcol < v1;
    lc.first = gcol == 0;
    lc.last = gcol + 2 >= LW;
    lc.seamL = XW && (lr & 31) == 0 && lr != 0;
    lc.seamR = XW && (lr & 31) == 31 && !lc.last;
    lc.post0 = XW && live && (lr & 31) == 0;
    lc.post31 = XW && live && (lr & 31) == 31;

    {
        const int n4 = (int)(lay.total() / 4);
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += NT) st4(smem + 4 * i, z4);
        __syncthreads();
    }

    // ---- per-graph scalars
    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const bool has_skip = a.p.skip != nullptr;
    const float s0 = has_skip ? a.p.skip[0] : 0.f, s1 = has_skip ? a.p.skip[1] : 1.f;
    const glrgtv_opparams& opT = COARSE ? a.p.gtv1 : a.p.gtv0;
    const glrgtv_opparams& opL = COARSE ? a.p.glr1 : a.p.glr0;
    const StatsTaps kT = glr_load_taps(opT.stats, c);
    const StatsTaps kL = HAS_L ? glr_load_taps(opL.stats, c) : kT;
    const float aT = expf(COARSE ? a.p.ro1[g] : a.p.ro0[g]);
    const float aL = HAS_L ? expf(COARSE ? a.p.mu1[g] : a.p.mu0[g]) : 0.f;
    const float Gam = THR ? expf(COARSE ? a.p.gamma1[g] : a.p.gamma0[g]) : 0.f;

    // ---- cp.async loader (bw2.cuh: unconditional copies, rows outside the image clamped for z / zero-filled for the weights).
    //      Fine: the lanes of a 16-byte piece share the copies (even lane: z, operand 1; odd lane: operands 0 and 2).
    int tl = R0 - 3;
    unsigned zo = 0;
    float* zdst = smem;
    bool zok = false;
    if (COARSE) {
        zok = live && 2 * x0 + 4 * lr < W;
        zo = (unsigned)(pl_f * a.s.H * W) + (unsigned)(2 * glr_clampi(tl, 0, LH - 1) * W + (zok ? 2 * x0 + 4 * lr : 0));
        zdst = smem + lay.zstage() + ((size_t)chc * B2_ZRC * 4 << LG) + 4 * lr;
    } else {
        const int cc = 2 * (lr & ~1);
        zok = live && x0 + cc < W;
        zo = (unsigned)(pl_f * a.s.H * W) + (unsigned)(glr_clampi(tl, 0, LH - 1) * W + (zok ? x0 + cc : 0));
        zdst = smem + lay.zsig() + ((size_t)chc * B2_ZR << LG) + cc;
    }
    float* const opdst = smem + lay.opring() + ((size_t)chc * B2_OPR << LG) + 2 * (lr & ~1);
    unsigned ooff = (unsigned)(pl_f * a.s.H * W) + (unsigned)((tl - 4) * W + x0 + 2 * (lr & ~1));
    float* const vcdst = smem + lay.vcring() + ((size_t)chc * B2_OPR << (LG - 1)) + (lr & ~3);
    unsigned vcoff = (unsigned)(pl_f * (LH / 2) * (LW / 2)) + (unsigned)(((tl - 4) >> 1) * (LW / 2) + x0 / 2 + (lr & ~3));
    const size_t vcsz = (size_t)a.s.B * G * F * (LH / 2) * (LW / 2);
    const float* wbase[3] = {a.cT, a.cT, a.cT};
    unsigned wo[3] = {0, 0, 0};
    float* wdst[3] = {smem, smem, smem};
    bool wok[3] = {false, false, false};
    int wlead[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {                      // up to three pieces per thread (6 NCH >= planes)
        const int wi = tid + j * NT;
        const int pl = wi >> (LG - 2), piece = wi & ((L >> 1) - 1);
        if (pl < NPLW && x0 + 4 * piece < LW) {
            const bool isL = HAS_L && pl >= PL_L, isRaw = THR && pl >= PL_RAW && pl < PL_L;
            const int e = isL ? pl - PL_L : isRaw ? pl - PL_RAW : pl;
            wlead[j] = (isL || isRaw) && e == 0 ? 1 : 0;
            wbase[j] = isL ? a.wL : isRaw ? a.wT : a.cT;
            wo[j] = (unsigned)((plane * ((isL || isRaw) ? 4 : 2) + e) * LH * LW) + (unsigned)(glr_clampi(tl - 2 + wlead[j], 0, LH - 1) * LW + x0 + 4 * piece);
            wdst[j] = smem + lay.wring() + ((size_t)pl * B2_WR << LG) + 4 * piece;
            wok[j] = true;
        }
    }
    auto issue = [&]() {
        if (zok) {
            if (COARSE) {
                float* d = zdst + ((tl & (B2_ZRC - 1)) * 4 << LG);
                cp_async16(d, a.z + zo); cp_async16(d + RW * 2, a.z + zo + W);
            } else if (!(lr & 1)) {
                cp_async16(zdst + ((tl & (B2_ZR - 1)) << LG), a.z + zo);
            }
        }
        if (tl >= 0 && tl < LH - 1) zo += COARSE ? 2 * W : W;
        if (NOP > 0 && zok && tl - 4 >= R0 && tl - 4 < R1) {
            float* d = opdst + (((tl - 4) & (B2_OPR - 1)) << LG);
            if (lr & 1) {
                cp_async16(d, a.op0 + ooff);
                if (NOP > 2 && a.op2 != nullptr) cp_async16(d + (2 * NCH * B2_OPR << LG), a.op2 + ooff);
            } else if (NOP > 1) {
                cp_async16(d + (NCH * B2_OPR << LG), a.op1 + ooff);
            }
        }
        ooff += W;
        if (!COARSE && live && (lr & 3) == 3 && x0 / 2 + (lr & ~3) < LW / 2 && !((tl - 4) & 1) && tl - 4 >= R0 - 1 && tl - 4 < R1) {
            float* d = vcdst + ((((tl - 4) >> 1) & (B2_OPR - 1)) << (LG - 1));
            cp_async16(d, a.vc_in + vcoff);
            if (NVC > 1) cp_async16(d + (NCH * B2_OPR << (LG - 1)), a.vc_in + vcsz + vcoff);
        }
        if ((tl - 4) & 1) vcoff += LW / 2;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int rw_ = tl - 2 + wlead[j];
            if (wok[j]) b2_cp16(wdst[j] + ((rw_ & (B2_WR - 1)) << LG), wbase[j] + wo[j], (unsigned)rw_ < (unsigned)LH);
            if (rw_ >= 0 && rw_ < LH - 1) wo[j] += LW;
        }
        ++tl;
    };

    // ---- walker state
    float2 sT[3], sL[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sT[k] = sL[k] = pzero();
    float2 oT = pzero(), oL = pzero(), oH = pzero();      // core outputs of the previous step (row t-3 at step t), already scaled by a_T / a_L
    float2 Fpend = pzero(), Fnext = pzero();              // scatter form of St: the linear part (GTV + GLR)
    float2 Hpend = pzero(), Hnext = pzero();              // X2: the thresholded GTV part
    float2 cDp = pzero(), wDpT = pzero();
    constexpr int N = 2, C = 1, U = 0;

    int t = R0 - 3;
    int zt = (t & (B2_ZR - 1)) << LG;
    int wt = ((t - 2) & (B2_WR - 1)) << LG;
    const int MBS = NCH * NWR * F2_NFLD * 2;
    float* const mb0 = smem + lay.mbox() + (chc * NWR + (lr >> 5)) * F2_NFLD * 2;
    int mofs = 0;
    unsigned goff = (unsigned)(pl_f * a.s.H * W) + (unsigned)((t - 4) * W + gcol);
    unsigned loff = (unsigned)(pl_f * LH * LW) + (unsigned)((t - 4) * LW + gcol);

#pragma unroll
    for (int k = 0; k < PD; ++k) { issue(); cp_async_commit(); }

#pragma unroll 1
    for (int m = 0; m < M; ++m) {
        cp_async_wait_pending<PD - 1>();
        __syncthreads();
        issue();
        cp_async_commit();
        const int o1 = (zt - RW) & M8, o2 = (zt - 2 * RW) & M8, o4 = (zt - 4 * RW) & M8;
        float* const mbw = mb0 + mofs;
        const float* const mbr = mb0 + (MBS - mofs);
        const int rf = t - 4;
        const bool fin = valid && rf >= R0 && rf < R1;

        // ---- COARSE: pool the two staged full-resolution rows of row t
        if (COARSE && live) {
            const float* zs = zdst + ((t & (B2_ZRC - 1)) * 4 << LG);
            float za[4], zb[4];
            ld4(zs, za); ld4(zs + 2 * RW, zb);
            pst(zsig + zt, make_float2(0.25f * (za[0] + za[1] + zb[0] + zb[1]), 0.25f * (za[2] + za[3] + zb[2] + zb[3])));
        }

        // ---- first-level stencils at row t-1 (s clamp-extended)
        {
            const float2 z0 = pld(zsig + zt), z1 = pld(zsig + o1), z2 = pld(zsig + o2);
            // (the first / last lane of a column strip's window owns halo columns: it takes its own value instead of the ring word
            // next door, which for channel 0 / slot 0 would be the word before the shared-memory block - tools/emu_asan.sh)
            const float2 z1l = pshl(z1, (lc.first || col0 == 0) ? z1.x : zsig[o1 - 1]), z1r = pshr(z1, (lc.last || col0 + 2 >= RW) ? z1.y : zsig[o1 + 2]);
            sT[N] = b2_S(kT, z1, z2, z0, z1l, z1r);
            if (HAS_L) sL[N] = b2_S(kL, z1, z2, z0, z1l, z1r);
            const int r = t - 1;
            if (r <= 0 || r >= LH) {
                if (r == 0) { sT[C] = sT[N]; sL[C] = sL[N]; }
                else if (r >= LH) { sT[N] = sT[C]; sL[N] = sL[C]; }
            }
            b2_post<XW>(sT[N], lc, mbw, F2_ST);
            if (HAS_L) b2_post<XW>(sL[N], lc, mbw, F2_SL);
        }

        // ---- St in scatter form, fed with the core outputs of row t-3 (computed one step ago): finishes row t-4
        float2 Fdone, Hdone = pzero();
        {
            float l, rr;
            b2_nb<true, XW>(oT, l, rr, lc, mbr, F2_OT);
            Fdone = pfmas(oT, kT.ku, Fpend);
            float2 np = pfmas(oT, kT.kc, Fnext);
            np = pfmas(pshl(oT, l), kT.kr, np);
            np = pfmas(pshr(oT, rr), kT.kl, np);
            float2 nn = pmuls(oT, kT.kd);
            if (HAS_L) {
                b2_nb<true, XW>(oL, l, rr, lc, mbr, F2_OL);
                Fdone = pfmas(oL, kL.ku, Fdone);
                np = pfmas(oL, kL.kc, np);
                np = pfmas(pshl(oL, l), kL.kr, np);
                np = pfmas(pshr(oL, rr), kL.kl, np);
                nn = pfmas(oL, kL.kd, nn);
            }
            Fpend = np; Fnext = nn;
            if (THR) {
                b2_nb<true, XW>(oH, l, rr, lc, mbr, F2_OH);
                Hdone = pfmas(oH, kT.ku, Hpend);
                float2 hp = pfmas(oH, kT.kc, Hnext);
                hp = pfmas(pshl(oH, l), kT.kr, hp);
                hp = pfmas(pshr(oH, rr), kT.kl, hp);
                Hpend = hp; Hnext = pmuls(oH, kT.kd);
            }
        }

        // ---- cores at row t-2
        {
            const int r = t - 2;
            const bool inimg = r >= 0 && r < LH;
            float sl, sr;
            b2_nb<false, XW>(sT[C], sl, sr, lc, mbr, F2_ST);
            const float2 dU = psub(sT[C], sT[U]), dL = psub(sT[C], pshl(sT[C], sl)), dR = psub(sT[C], pshr(sT[C], sr)), dD = psub(sT[C], sT[N]);
            {
                // linear core Ct C with the symmetric coefficients cR, cD (self-adjoint); cD of the row above is carried
                const float* crow = wring + wt;
                const float2 cr = pld(crow), cd = pld(crow + (B2_WR << LG));
                const float2 crl = make_float2(lc.first ? 0.f : crow[-1], cr.x);
                float2 o = pmul(cDp, dU);
                o = pfma(crl, dL, o); o = pfma(cr, dR, o); o = pfma(cd, dD, o);
                cDp = cd;
                oT = inimg ? pmuls(o, aT) : pzero();
            }
            if (THR) {
                const float* w0 = wring + ((size_t)PL_RAW * B2_WR << LG) + wt;
                const float* pL = w0 + (1 * B2_WR << LG);
                const float* pR = w0 + (2 * B2_WR << LG);
                float2 own[4], in[4];
                own[0] = pld(w0); own[1] = pld(pL); own[2] = pld(pR); own[3] = pld(w0 + (3 * B2_WR << LG));
                in[0] = wDpT;
                in[3] = pld(wring + ((size_t)PL_RAW * B2_WR << LG) + ((wt + RW) & M4));
                in[1] = make_float2(lc.first ? 0.f : pR[-1], own[2].x);
                in[2] = make_float2(own[1].y, lc.last ? 0.f : pL[2]);
                wDpT = own[3];
                const float2 dd[4] = {dU, dL, dR, dD};
                float ox = 0.f, oy = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ox += own[e].x * glr_phi(own[e].x * dd[e].x, Gam) + in[e].x * glr_phi(in[e].x * dd[e].x, Gam);
                    oy += own[e].y * glr_phi(own[e].y * dd[e].y, Gam) + in[e].y * glr_phi(in[e].y * dd[e].y, Gam);
                }
                oH = inimg ? make_float2(aT * ox, aT * oy) : pzero();
            }
            if (HAS_L) {
                b2_nb<false, XW>(sL[C], sl, sr, lc, mbr, F2_SL);
                const float* w0 = wring + ((size_t)PL_L * B2_WR << LG) + wt;
                const float2 wU = pld(w0), wLe = pld(w0 + (1 * B2_WR << LG)), wRi = pld(w0 + (2 * B2_WR << LG)), wD = pld(w0 + (3 * B2_WR << LG));
                float2 acc = pmul(wU, sL[U]);
                acc = pfma(wLe, pshl(sL[C], sl), acc); acc = pfma(wRi, pshr(sL[C], sr), acc); acc = pfma(wD, sL[N], acc);
                oL = inimg ? pmuls(psub(sL[C], acc), aL) : pzero();
            }
            b2_post<XW>(oT, lc, mbw, F2_OT);
            if (HAS_L) b2_post<XW>(oL, lc, mbw, F2_OL);
            if (THR) b2_post<XW>(oH, lc, mbw, F2_OH);
        }

        // ---- epilogue of row t-4
        if (fin) {
            if (COARSE) {
                pst(a.vc_out + loff, Fdone);
                if (THR) pst(a.vc_out + (size_t)a.s.B * G * F * LH * LW + loff, Hdone);
            } else {
                const float2 zq = pld(zsig + o4);
                const float* vcs = vcdst + (((rf >> 1) & (B2_OPR - 1)) << (LG - 1)) + (lr & 3);
                const float2 Rl = pfmas(pset(vcs[0]), 0.25f, Fdone);                       // (A - I) z : linear GTV + GLR, both resolutions
                const float* ops = opdst - 2 * (lr & ~1) + col0 + ((rf & (B2_OPR - 1)) << LG);
                if (MODE == FW_BA) {
                    pst(a.out0 + goff, padd(zq, Rl));
                } else if (MODE == FW_X1) {
                    pst(a.out0 + goff, pfmas(Rl, -al0, zq));                                // x1 = bA + a0 (bA - A bA)
                } else if (MODE == FW_X2) {                                                 // ops: y
                    const float2 Rh = pfmas(pset(vcs[NCH * B2_OPR << (LG - 1)]), 0.25f, Hdone);
                    const float2 bB = padd(pld(ops), Rh);
                    const float2 r1 = psub(bB, padd(zq, Rl));
                    pst(a.out1 + goff, bB);
                    pst(a.out2 + goff, r1);
                    pst(a.out0 + goff, pfmas(r1, al1, zq));
                } else {                                                                    // X3: ops bB, r1, x
                    const float2 bB = pld(ops), r1 = pld(ops + (NCH * B2_OPR << LG));
                    const float2 u2 = pfmas(r1, be2, psub(bB, padd(zq, Rl)));
                    float2 x3 = pfmas(u2, al2, zq);
                    if (has_skip) x3 = pfmas(pld(ops + (2 * NCH * B2_OPR << LG)), s0, pmuls(x3, s1));
                    pst(a.out0 + goff, x3);
                }
            }
        }
        sT[U] = sT[C]; sT[C] = sT[N];
        if (HAS_L) { sL[U] = sL[C]; sL[C] = sL[N]; }
        ++t;
        zt = (zt + RW) & M8;
        wt = (wt + RW) & M4;
        goff += W; loff += LW;
        mofs = MBS - mofs;
    }
    cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------
// host side (templates: instantiated per stage in fw2_*.cu)
// ---------------------------------------------------------------------------------------------------
// window of one CTA: all of a plane up to 2 * B2_MAXT / (channels) ... columns; wider planes are cut into strips
static inline int f2_lanes(int lw, int F) {
    int L = b2_lanes(lw);
    while (L > 4 && b2_nch(F, L) * 2 < F) L /= 2;           // a CTA holds at least half of a graph's channels
    return L;
}
template <int MODE, bool COARSE>
static bool f2_fits(const glrgtv_shape* s) {
    const int L = f2_lanes(COARSE ? s->W / 2 : s->W, s->F), nch = b2_nch(s->F, L);
    if (L < 8 || nch < 1 || 6 * nch < F2Smem<MODE, COARSE>::NPLW) return false;
    F2Smem<MODE, COARSE> lay; lay.NCH = nch; lay.L = L;
    return lay.bytes() + 256 <= 227 * 1024;
}

template <int MODE, bool COARSE, int LGT, int NCHT, bool XWG>
static int f2_go(const F2Args& a, long blocks, int threads, size_t smem, void* stream) {
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_fw2<MODE, COARSE, LGT, NCHT, XWG>, smem, optin)) return rc_;
#endif
    GLR_LAUNCH_FIBERS((k_fw2<MODE, COARSE, LGT, NCHT, XWG>), dim3((unsigned)blocks), threads, smem, stream, a);
    return GLRGTV_OK;
}

template <int MODE, bool COARSE>
static int f2_launch(F2Args a, void* stream) {
    const glrgtv_shape& s = a.s;
    const int LW = COARSE ? s.W / 2 : s.W;
    const int L = f2_lanes(LW, s.F);
    int lg = 0;
    while ((1 << lg) < 2 * L) ++lg;
    a.lg = lg;
    a.nch = b2_nch(s.F, L);
    if (a.nch < 1) return GLRGTV_ERR_UNSUPPORTED;
    a.n_parts = s.F / a.nch;
    if (2 * L >= LW) { a.n_strips = 1; a.strip_w = LW; }
    else { a.strip_w = 2 * L - 2 * F2_HALO; a.n_strips = (LW + a.strip_w - 1) / a.strip_w; }
    const int threads = (a.nch * L + 31) & ~31;
    F2Smem<MODE, COARSE> lay; lay.NCH = a.nch; lay.L = L;
    const size_t smem = lay.bytes();
    if (threads > B2_MAXT || smem > 227 * 1024) return GLRGTV_ERR_UNSUPPORTED;
    const int rows = a.row1 - a.row0;
    int occ = (int)(227 * 1024 / (smem + 1024));
    const int regocc = (MODE == FW_X2 ? 1 : 2) * B2_MAXT / threads;      // the light stages are built for two CTAs of B2_MAXT threads per SM
    if (occ > regocc) occ = regocc;
    if (occ < 1) occ = 1;
    int bands = 1;
    {
        const long base = (long)s.B * s.G * a.n_parts * a.n_strips, slots = 148L * occ;
        long best = -1;
        for (int bnd = 1; bnd <= 16 && rows / bnd >= 24; bnd *= 2) {
            const long cost = ((base * bnd + slots - 1) / slots) * ((rows + bnd - 1) / bnd + 7);
            if (best < 0 || cost < best) { best = cost; bands = bnd; }
        }
    }
    a.band_rows = (rows + bands - 1) / bands;
    a.n_bands = (rows + a.band_rows - 1) / a.band_rows;
    const long blocks = (long)s.B * s.G * a.n_parts * a.n_bands * a.n_strips;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    ++g_glr_stream_launches;
    int rc = -1000;
#ifndef GLRGTV_EMU
#define F2_TRY(LG_, NCH_) if (rc == -1000 && lg == LG_ && a.nch == NCH_) rc = f2_go<MODE, COARSE, LG_, NCH_, false>(a, blocks, threads, smem, stream);
    F2_TRY(8, 3) F2_TRY(7, 6) F2_TRY(6, 12) F2_TRY(6, 6) F2_TRY(5, 12) F2_TRY(4, 12)
#undef F2_TRY
#endif
    if (rc == -1000)
        rc = L > 32 ? f2_go<MODE, COARSE, 0, 0, true>(a, blocks, threads, smem, stream) : f2_go<MODE, COARSE, 0, 0, false>(a, blocks, threads, smem, stream);
    return rc ? rc : GLR_CHECK_LAUNCH();
}

// one forward stage on the level rows [row0, row1) (full-resolution rows; even bounds): the half-resolution launch, then the
// full-resolution one.  `vc`: scratch [nvc][B,C,H/2,W/2].
template <int MODE>
int glr_fw2_stage(F2Args a, const float* cT1, const float* wT1, const float* wL1, float* vc, int row0, int row1, void* stream) {
    F2Args c = a;
    c.cT = cT1; c.wT = wT1; c.wL = wL1; c.vc_out = vc; c.vc_in = nullptr;
    c.row0 = row0 / 2; c.row1 = row1 / 2;
    int rc = f2_launch<MODE, true>(c, stream);
    if (rc) return rc;
    a.vc_in = vc; a.vc_out = nullptr; a.row0 = row0; a.row1 = row1;
    return f2_launch<MODE, false>(a, stream);
}
