// block_stream_fwd.cu - register-streaming forward stages of LocalLowpassFilteringBlock / MixtureGTVGLR
// (V1X0:707-811, 985-988); same stage cut as block_fwd.cu (BA, X1, X2, X3), different execution model (stream.cuh).
//
// One CTA = one (batch, graph), `nch` of its channels, a band of rows.  Per channel there is a FINE walker (GL lanes,
// the full-resolution chain and the stage epilogue) and a COARSE walker (GL/2 lanes, the half-resolution chain on the
// 2x2 mean of the input).  Both are instantiations of ONE step template (stream_walk); they differ in how a row is
// loaded (direct / pooled) and where the finished row goes (global memory / a two-slot shared-memory ring that hands
// the coarse term 0.25 * P^T[...] to the fine epilogue).  The coarse walker steps on even block steps only and runs
// DF = 8 steps ahead of the fine walker, which is exactly the depth of its pipeline in fine rows.  A launch may be
// restricted to a row range of the plane (glrgtv_block_fwd_stage): walkers then start mid-image on real rows.
//
// Eligible shapes: W % 8 == 0.  A walker is at most two warps (256 columns) wide; wider planes (4K inference) are cut
// into column strips of STREAM_STRIP valid columns with 8 halo columns per side - the reach of the chain through the
// half-resolution branch - whose edge lanes compute throw-away values.  Other widths take the plane kernels of block_fwd.cu.
#include "stream.cuh"
#include "fw2.cuh"

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
    float* vc;                // scratch of the second-generation kernels (fw2.cuh), or NULL
    float* out0;  // BA: bA | X1: x1 | X2: x2 | X3: out
    float* out1;  // X2: bB
    float* out2;  // X2: r1
    int nch;        // channels per CTA (divides F)
    int band_rows;  // fine rows per CTA (even)
    int n_bands;
    int row_base, row_end;   // the rows this launch produces: [row_base, row_end) (even bounds; the whole plane by default)
    int n_strips;   // column strips (planes wider than a 64-lane walker): STREAM_STRIP valid columns each, 8 columns of halo per side
};

// one resolution of one walker
struct Lvl {
    int H, W;
    StatsTaps kL, kT;
    float aL, aT, Gam;
};

// Shared-memory staging.  Every global operand reaches the walkers through rings that ONE elected thread fills with
// TMA bulk copies (cp.async.bulk, whole image rows, completion counted on an mbarrier) two to three block steps ahead of
// their use, so no walker ever issues or waits on a global load of its own:
//   zring  [nch][ZR][Wp]          stage-input rows; the coarse walker meets them first (pooling two per coarse row), the
//                                 fine walker reads them again DF steps later: z is loaded once
//   opring [nch][NOP][OPR][Wp]    epilogue operands (X2: y | X3: x, bB, r1)
//   w0ring [NPL][WR][Wp], w1ring [NPL][WR][Wp/2]   weight / coefficient rows of this (batch, graph), ONE copy per CTA
//                                 shared by its channel walkers; plane order: cT(2), wL(4) if GLR, wT(4) if THR.
//                                 The wT plane U is staged one row ahead (its row r+1 feeds the thresholded core).
//   cring  [nch][2][NRING][Wp/2]  coarse result -> fine epilogue hand-over
// Copies are issued in BATCHES: batch j = everything block steps 2j-1 and 2j read, issued at the odd step 2j-3 by lane 0
// of the first coarse warp (coarse walkers rest on odd steps), tracked by mbarrier j % 4.
#ifndef STREAM_MAXT
#define STREAM_MAXT 192   // threads per CTA (fine + coarse walkers)
#endif
#ifndef STREAM_MINB
#define STREAM_MINB 2     // resident CTAs per SM the register budget is sized for
#endif
#define STREAM_ZR 14
#define STREAM_OPR 4
#define STREAM_WR 4
#define STREAM_NBAR 4
#define STREAM_STRIP 240   // valid output columns of one strip of a wide plane (256-column walker minus 2 x 8 halo columns)
#ifndef STREAM_PHASES
#define STREAM_PHASES 0   // 0: per stage (see stream_walk), 1: rotated windows everywhere, 3: phase-unrolled everywhere
#endif
#define STREAM_PD 2      // cp.async loader: block steps between issuing a copy and reading it
#define STREAM_MAXJ 5    // cp.async loader: weight-plane copies per thread and step held as precomputed descriptors
template <int MODE>
struct StreamSmem {
    static constexpr bool GLR = MODE != MODE_BA, THR = MODE == MODE_X2;
    static constexpr int NOP = MODE == MODE_X2 ? 1 : MODE == MODE_X3 ? 3 : 0;
    static constexpr int NPL = 2 + (GLR ? 4 : 0) + (THR ? 4 : 0);
    static constexpr int NRING = THR ? 2 : 1;
    static constexpr int PL_C = 0, PL_WL = 2, PL_WT = 2 + (GLR ? 4 : 0);
    int Wp, nch;
    __host__ __device__ size_t bars() const { return 0; }                       // STREAM_NBAR x 8 bytes
    __host__ __device__ size_t zring() const { return 2 * STREAM_NBAR; }
    __host__ __device__ size_t opring() const { return zring() + (size_t)nch * STREAM_ZR * Wp; }
    __host__ __device__ size_t w0ring() const { return opring() + (size_t)nch * NOP * STREAM_OPR * Wp; }
    __host__ __device__ size_t w1ring() const { return w0ring() + (size_t)NPL * STREAM_WR * Wp; }
    __host__ __device__ size_t cring() const { return w1ring() + (size_t)NPL * STREAM_WR * (Wp / 2); }
    __host__ __device__ size_t mbox() const { return cring() + (size_t)nch * 2 * NRING * (Wp / 2); }
    __host__ __device__ size_t total() const { return mbox() + (size_t)2 * nch * PL_COUNT * 2 + 8; }
};

// geometry of one CTA, shared by the walkers and the producer
struct StreamCta {
    int H, W, F, G, nch, b, g, f0, R0, R1, K0, M, Wp;
    int x0, v0, v1;      // first column of the walker's window; valid output columns [v0, v1)
    size_t HW;
};
__device__ __forceinline__ StreamCta stream_cta(const StreamFwdArgs& a, int GL) {
    StreamCta c;
    c.H = a.s.H; c.W = a.s.W; c.F = a.s.F; c.G = a.s.G; c.nch = a.nch;
    int bid = (int)blockIdx.x;
    const int strip = bid % a.n_strips; bid /= a.n_strips;
    const int band = bid % a.n_bands; bid /= a.n_bands;
    const int chunks = c.F / c.nch;
    const int chunk = bid % chunks; bid /= chunks;
    c.g = bid % c.G; c.b = bid / c.G;
    c.f0 = chunk * c.nch;
    c.R0 = a.row_base + band * a.band_rows;
    c.R1 = c.R0 + a.band_rows < a.row_end ? c.R0 + a.band_rows : a.row_end;
    c.K0 = c.R0 / 2;
    c.M = (c.R1 - c.R0) + 6 + STREAM_DF;
    if (a.n_strips > 1) {
        c.v0 = strip * STREAM_STRIP;
        c.v1 = c.v0 + STREAM_STRIP < c.W ? c.v0 + STREAM_STRIP : c.W;
        c.x0 = strip ? c.v0 - 8 : 0;
    } else {
        c.x0 = 0; c.v0 = 0; c.v1 = c.W;
    }
    c.Wp = 4 * GL;
    c.HW = (size_t)c.H * c.W;
    return c;
}

// ---- the producer: all global -> shared traffic of batch j (block steps 2j-1 and 2j), as TMA bulk row copies.
// The 32 lanes of the first coarse warp share the batch's item list
//     [fine items of step 2j-1 | fine items of step 2j | coarse items of step 2j],
//     fine items = NPL weight rows + nch*NOP operand rows, coarse items = NPL weight rows + 2*nch input rows;
// lane l owns items l, l+32, ...  An item's row is an affine function of j, so each lane decodes its items ONCE into
// descriptors and a batch costs it a handful of instructions per copy.
#define STREAM_MAXI 4
struct ProdItem {
    const float* src0;   // row 0 of the source plane (nullptr: no item)
    int dst0;            // ring base, floats from the start of shared memory
    int r_off, rstep;    // row = r_off + rstep * j
    int lo, hi;          // copy iff lo <= row < hi
    int pitch;           // floats per source row (= bytes / 4 of one copy)
    int rpitch;          // ring row pitch in floats; < 0: the z ring (slot = (2j mod ZR) + k, pitch -rpitch)
};
template <int MODE>
__device__ __forceinline__ ProdItem stream_item(const StreamFwdArgs& a, const StreamCta& c, int it) {
    using SM = StreamSmem<MODE>;
    constexpr int NOP = SM::NOP, NPL = SM::NPL, ZR = STREAM_ZR;
    SM lay; lay.Wp = c.Wp; lay.nch = c.nch;
    const int Wp = c.Wp, Wpc = Wp / 2, H = c.H, W = c.W, Hc = H / 2, Wc = W / 2;
    const size_t HW = c.HW, HWc = HW / 4, plane = (size_t)c.b * c.G + c.g;
    const size_t off0 = ((size_t)c.b * c.G * c.F + (size_t)c.g * c.F + c.f0) * HW;     // first channel of this CTA
    const bool has_skip = MODE == MODE_X3 && a.p.skip != nullptr;
    const int NF = NPL + c.nch * NOP, NC = NPL + 2 * c.nch, NI = 2 * NF + NC;
    ProdItem p;
    p.src0 = nullptr; p.dst0 = 0; p.r_off = 0; p.rstep = 0; p.lo = 0; p.hi = 0; p.pitch = 0; p.rpitch = 0;
    if (it >= NI) return p;
    const bool coarse = it >= 2 * NF;
    const int s = coarse ? 1 : (it >= NF ? 1 : 0);                  // block step mt = 2j - 1 + s
    int idx = coarse ? it - 2 * NF : (it >= NF ? it - NF : it);
    // newest row of the consuming walker at step mt: fine t = R0-3 + mt-DF, coarse t = K0-3 + mt/2; mt in [DF or 0, M)
    const int t_off = coarse ? c.K0 - 3 : c.R0 - 4 - STREAM_DF + s;    // t = t_off + (coarse ? j : 2j)
    const int t_lo = coarse ? c.K0 - 3 : c.R0 - 3;                     // t at the first step that consumes
    const int t_hi = coarse ? c.K0 - 3 + ((c.M + 1) >> 1) : c.R0 - 3 + c.M - STREAM_DF;   // one past the last
    p.rstep = coarse ? 1 : 2;
    if (idx < NPL) {                                     // weight / coefficient row of the core row t-2
        const int set = idx < 2 ? 0 : (SM::GLR && idx < 6) ? 1 : 2;
        const int e = idx - (set == 0 ? 0 : set == 1 ? SM::PL_WL : SM::PL_WT);
        const int d = -2 + ((set == 2 && e == 0) ? 1 : 0);
        p.r_off = t_off + d;
        p.lo = glr_maxi(0, t_lo + d); p.hi = glr_mini(coarse ? Hc : H, t_hi + d);
        if (coarse) {
            p.src0 = (set == 0 ? a.cT1 + plane * 2 * HWc : (set == 1 ? a.wL1 : a.wT1) + plane * 4 * HWc) + (size_t)e * HWc;
            p.dst0 = (int)lay.w1ring() + idx * STREAM_WR * Wpc; p.pitch = Wc; p.rpitch = Wpc;
        } else {
            p.src0 = (set == 0 ? a.cT0 + plane * 2 * HW : (set == 1 ? a.wL0 : a.wT0) + plane * 4 * HW) + (size_t)e * HW;
            p.dst0 = (int)lay.w0ring() + idx * STREAM_WR * Wp; p.pitch = W; p.rpitch = Wp;
        }
        return p;
    }
    idx -= NPL;
    if (coarse) {                                        // the two fine rows coarse row t pools, per channel: rho = 2t + k
        const int ch = idx >> 1, k = idx & 1;
        p.src0 = a.z + off0 + (size_t)ch * HW;
        p.r_off = 2 * t_off + k; p.rstep = 2;
        p.lo = glr_maxi(0, 2 * t_lo + k); p.hi = glr_mini(H, 2 * (t_hi - 1) + k + 1);
        p.dst0 = (int)lay.zring() + (ch * ZR + k) * Wp; p.pitch = W; p.rpitch = -Wp;
        return p;
    }
    if (NOP > 0) {                                       // epilogue operands of row t-3, per channel
        const int ch = idx / (NOP > 0 ? NOP : 1), k = idx - ch * NOP;
        if (MODE == MODE_X3 && k == 0 && !has_skip) return p;
        p.src0 = (k == 0 ? a.y : k == 1 ? a.bB_in : a.r1_in) + off0 + (size_t)ch * HW;
        p.r_off = t_off - 3;
        p.lo = glr_maxi(c.R0, t_lo - 3); p.hi = glr_mini(c.R1, t_hi - 3);
        p.dst0 = (int)lay.opring() + (ch * NOP + k) * STREAM_OPR * Wp; p.pitch = W; p.rpitch = Wp;
    }
    return p;
}
// issue batch j: called by all 32 producer lanes
__device__ __forceinline__ void stream_produce(const ProdItem (&items)[STREAM_MAXI], float* smem, int bars_off, int j, int pl_lane) {
    const smem_addr_t sbase = smem_addr(smem);
    const smem_addr_t bar = smem_advance(sbase, bars_off + 2 * (j & (STREAM_NBAR - 1)));
    int row[STREAM_MAXI];
    unsigned bytes = 0;
#pragma unroll
    for (int i = 0; i < STREAM_MAXI; ++i) {
        row[i] = items[i].r_off + items[i].rstep * j;
        if (items[i].src0 != nullptr && row[i] >= items[i].lo && row[i] < items[i].hi) bytes += 4u * (unsigned)items[i].pitch;
        else row[i] = -1;
    }
    // total over the warp (exact in fp32 far beyond any batch size), then arm the barrier before the copies fly
    float fb = (float)bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fb += __shfl_xor_sync(0xffffffffu, fb, o);
    if (pl_lane == 0) mbar_expect_tx(bar, (unsigned)fb);
    __syncwarp();
    const int zslot = (2 * j) % STREAM_ZR;
#pragma unroll
    for (int i = 0; i < STREAM_MAXI; ++i) {
        if (row[i] < 0) continue;
        const int slot_off = items[i].rpitch < 0 ? zslot * (-items[i].rpitch) : (row[i] & (STREAM_WR - 1)) * items[i].rpitch;
        bulk_g2s(smem_advance(sbase, items[i].dst0 + slot_off), items[i].src0 + (size_t)row[i] * items[i].pitch,
                 4u * (unsigned)items[i].pitch, bar);
    }
}

// The walk of one role.  FINE / COARSE are separate instantiations (their warps never mix), so each loop is
// specialised; both execute exactly M block barriers.  The row windows are 3-cycles indexed by the compile-time phase
// of the unrolled step, so no register is ever moved to "rotate" a window.
template <int MODE, bool XW, bool FINE, bool TMA>
__device__ __forceinline__ void stream_walk(const StreamFwdArgs& a, float* smem, const int wk, const int nwk, const int lane,
                                            const int GL) {
    using SM = StreamSmem<MODE>;
    constexpr bool GLR = SM::GLR, THR = SM::THR;
    constexpr int NRING = SM::NRING, NOP = SM::NOP, NPL = SM::NPL, ZR = STREAM_ZR, PD = STREAM_PD;
    const StreamCta ct = stream_cta(a, GL);
    const int H = ct.H, W = ct.W, G = ct.G, nch = ct.nch, g = ct.g, R0 = ct.R0, R1 = ct.R1, M = ct.M;
    const bool live = wk < nch;
    const int c = g * ct.F + ct.f0 + (live ? wk : 0);
    const size_t off = ((size_t)ct.b * G * ct.F + c) * ct.HW;

    SM lay; lay.Wp = ct.Wp; lay.nch = nch;
    const int Wp = ct.Wp, Wpc = Wp / 2;
    const float* zring = smem + lay.zring() + (size_t)(live ? wk : 0) * ZR * Wp;              // this walker's channel
    const float* opring = smem + lay.opring() + (size_t)(live ? wk : 0) * NOP * STREAM_OPR * Wp;
    const float* wring = smem + (FINE ? lay.w0ring() : lay.w1ring());                          // this role's level
    float* cring = smem + lay.cring() + (size_t)(live ? wk : 0) * 2 * NRING * Wpc;
    float* mbox = smem + lay.mbox();
    const smem_addr_t bars = smem_advance(smem_addr(smem), (int)lay.bars());
    const int wpitch = FINE ? Wp : Wpc;                                                  // row pitch of this level's weight ring
    const int GLc_ = GL / 2;
    const int pw_lane = wk * GLc_ + lane;                  // lane within the first coarse warp (valid when < 32)
    const bool producer = TMA && !FINE && pw_lane < 32;

    Lvl L;
    if (FINE) {
        L.H = H; L.W = W;
        L.kT = glr_load_taps(a.p.gtv0.stats, c);
        L.kL = GLR ? glr_load_taps(a.p.glr0.stats, c) : L.kT;
        L.aT = expf(a.p.ro0[g]); L.aL = GLR ? expf(a.p.mu0[g]) : 0.f; L.Gam = THR ? expf(a.p.gamma0[g]) : 0.f;
    } else {
        L.H = H / 2; L.W = W / 2;
        L.kT = glr_load_taps(a.p.gtv1.stats, c);
        L.kL = GLR ? glr_load_taps(a.p.glr1.stats, c) : L.kT;
        L.aT = expf(a.p.ro1[g]); L.aL = GLR ? expf(a.p.mu1[g]) : 0.f; L.Gam = THR ? expf(a.p.gamma1[g]) : 0.f;
    }

    LaneCtx lc;
    const int lcol = 4 * lane;                             // column inside the walker's window (shared-memory rings)
    lc.col0 = (FINE ? ct.x0 : ct.x0 / 2) + lcol;           // column in the image plane of this resolution
    lc.width = FINE ? (GL < 32 ? GL : 32) : (GL / 2 < 32 ? GL / 2 : 32);
    lc.active = live && lc.col0 < L.W;
    lc.first = lc.col0 == 0;
    lc.last = lc.col0 + 4 >= L.W;
    lc.seam_l = XW && FINE && live && lane == 32 && lc.col0 < L.W;
    lc.seam_r = XW && FINE && live && lane == 31 && lc.col0 + 4 < L.W;
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

    const int r0 = FINE ? R0 : ct.K0;
    auto wrow = [&](int pl, int row) { return wring + (pl * STREAM_WR + (row & (STREAM_WR - 1))) * wpitch + lcol; };

    // ---- cp.async loader (TMA == false): every thread stages its own share STREAM_PD block steps ahead.
    //      weight planes wk, wk+nwk, ... of this role's level (src at row 0, ring address, row lead), this lane's quad
    const smem_addr_t sbase = smem_addr(smem);
    const float* wsrc[STREAM_MAXJ];
    smem_addr_t wdst[STREAM_MAXJ];
    bool wlead[STREAM_MAXJ];
    const float* opp[3] = {nullptr, nullptr, nullptr};
    const float* zpp = a.z + off + 2 * lc.col0;
    int zis = 0;                                           // ring slot of the next pooled row pair to stage
    if (!TMA) {
        const size_t plane = (size_t)ct.b * G + g;
        const int LHW = L.H * L.W;
        const bool copier = lc.col0 < L.W;                 // dead walkers help with the shared weight copies
#pragma unroll
        for (int j = 0; j < STREAM_MAXJ; ++j) {
            const int pl = wk + j * nwk;
            wsrc[j] = nullptr; wdst[j] = sbase; wlead[j] = false;
            if (copier && pl < NPL) {
                const int set = pl < 2 ? 0 : (GLR && pl < 6) ? 1 : 2;
                const int e = pl - (set == 0 ? 0 : set == 1 ? SM::PL_WL : SM::PL_WT);
                const float* base = FINE ? (set == 0 ? a.cT0 + plane * 2 * LHW : (set == 1 ? a.wL0 : a.wT0) + plane * 4 * LHW)
                                         : (set == 0 ? a.cT1 + plane * 2 * LHW : (set == 1 ? a.wL1 : a.wT1) + plane * 4 * LHW);
                wsrc[j] = base + (size_t)e * LHW + lc.col0;
                wdst[j] = smem_advance(sbase, (int)(FINE ? lay.w0ring() : lay.w1ring()) + pl * STREAM_WR * wpitch + lcol);
                wlead[j] = set == 2 && e == 0;             // wT plane U runs one row ahead
            }
        }
        if (FINE && lc.active) {
            if (MODE == MODE_X2 || has_skip) opp[0] = a.y + off + lc.col0;
            if (MODE == MODE_X3) { opp[1] = a.bB_in + off + lc.col0; opp[2] = a.r1_in + off + lc.col0; }
        }
    }
    const smem_addr_t opdst = smem_advance(sbase, (int)lay.opring() + (live ? wk : 0) * NOP * STREAM_OPR * Wp + lcol);
    const smem_addr_t zdst = smem_advance(sbase, (int)lay.zring() + (live ? wk : 0) * ZR * Wp + 2 * lcol);
    auto issue = [&](int mt) {
        if (mt >= M) return;
        if (FINE ? (mt < STREAM_DF) : (mt & 1)) return;
        const int t = r0 - 3 + (FINE ? mt - STREAM_DF : (mt >> 1));
        if (!FINE) {
            // the two fine rows the coarse row t pools (this lane's 8 columns)
            if (lc.active) {
                const int rho = 2 * t;
                const smem_addr_t d = smem_advance(zdst, zis * Wp);
                const float* src = zpp + (size_t)rho * W;
                if (rho >= 0 && rho < H) { cp_async16_s(d, src); cp_async16_s(smem_advance(d, 4), src + 4); }
                if (rho + 1 >= 0 && rho + 1 < H) { cp_async16_s(smem_advance(d, Wp), src + W); cp_async16_s(smem_advance(d, Wp + 4), src + W + 4); }
            }
            zis = zis + 2 == ZR ? 0 : zis + 2;
        } else if (NOP > 0) {
            const int re = t - 3;                          // epilogue operands of row t-3
            if (re >= R0 && re < R1) {
                const int so = (re & (STREAM_OPR - 1)) * Wp, go = re * W;
#pragma unroll
                for (int k = 0; k < NOP; ++k)
                    if (opp[k] != nullptr) cp_async16_s(smem_advance(opdst, k * STREAM_OPR * Wp + so), opp[k] + go);
            }
        }
        // weight / coefficient rows of the core row t-2 of this level, shared by the CTA's channel walkers
        const int rw = t - 2;
        const bool ok0 = rw >= 0 && rw < L.H, ok1 = rw + 1 >= 0 && rw + 1 < L.H;
        const int so0 = (rw & (STREAM_WR - 1)) * wpitch, so1 = ((rw + 1) & (STREAM_WR - 1)) * wpitch;
        const int go0 = rw * L.W, go1 = go0 + L.W;
#pragma unroll
        for (int j = 0; j < STREAM_MAXJ; ++j) {
            if (wsrc[j] == nullptr) continue;
            if (wlead[j]) { if (ok1) cp_async16_s(smem_advance(wdst[j], so1), wsrc[j] + go1); }
            else if (ok0) cp_async16_s(smem_advance(wdst[j], so0), wsrc[j] + go0);
        }
        if (wk + STREAM_MAXJ * nwk < NPL && lc.col0 < L.W) {      // (a CTA with very few walkers: remaining planes, generic path)
            const size_t plane = (size_t)ct.b * G + g;
            const int LHW = L.H * L.W;
            for (int pl = wk + STREAM_MAXJ * nwk; pl < NPL; pl += nwk) {
                const int set = pl < 2 ? 0 : (GLR && pl < 6) ? 1 : 2;
                const int e = pl - (set == 0 ? 0 : set == 1 ? SM::PL_WL : SM::PL_WT);
                const bool lead = set == 2 && e == 0;
                if (lead ? !ok1 : !ok0) continue;
                const float* base = FINE ? (set == 0 ? a.cT0 + plane * 2 * LHW : (set == 1 ? a.wL0 : a.wT0) + plane * 4 * LHW)
                                         : (set == 0 ? a.cT1 + plane * 2 * LHW : (set == 1 ? a.wL1 : a.wT1) + plane * 4 * LHW);
                cp_async16_s(smem_advance(sbase, (int)(FINE ? lay.w0ring() : lay.w1ring()) + pl * STREAM_WR * wpitch + lcol + (lead ? so1 : so0)),
                             base + (size_t)e * LHW + lc.col0 + (lead ? go1 : go0));
            }
        }
    };

    Row z[3], sA[3], sB[3], lA[3], oB[3], oT[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) z[k] = sA[k] = sB[k] = lA[k] = oB[k] = oT[k] = row_zero();
    Row cDp = row_zero(), wU_next = row_zero(), wD_prev = row_zero();
    // ring slots of the rows this walker reads next: fine row t / fine row t-3 / first pooled row (row - (R0-6) mod ZR)
    int zs = FINE ? STREAM_DF - 5 : 0, zqs = 0;

    if (!TMA) {
#pragma unroll
        for (int k = 0; k < PD; ++k) { issue(k); cp_async_commit(); }
    }
    ProdItem items[STREAM_MAXI];
    if (TMA && !FINE) {
#pragma unroll
        for (int i = 0; i < STREAM_MAXI; ++i) items[i] = stream_item<MODE>(a, ct, producer ? pw_lane + 32 * i : 1 << 20);
        if (producer) { stream_produce(items, smem, (int)lay.bars(), 0, pw_lane); stream_produce(items, smem, (int)lay.bars(), 1, pw_lane); }
    }

    // PH3: the step is unrolled over the three window phases (no register moves, 6x the code); otherwise one step body
    // whose windows are rotated with register moves.  The 32 KB L1.5 instruction cache decides: measured on B200, the
    // thresholded stage X2 (the largest body) is faster rotated, the other three are faster unrolled.
    constexpr bool PH3 = STREAM_PHASES == 3 || (STREAM_PHASES == 0 && MODE != MODE_X2);
    constexpr int KU = PH3 ? 6 : 1;
#pragma unroll 1
    for (int m0 = 0; m0 < M; m0 += KU) {
#pragma unroll
        for (int ku = 0; ku < KU; ++ku) {
            const int m = m0 + ku;
            if (m >= M) break;
            const int k = PH3 ? ku : (m & 1);
            if (TMA) {
                const int j = (m + 1) >> 1;
                mbar_wait(smem_advance(bars, 2 * (j & (STREAM_NBAR - 1))), (unsigned)(j >> 2) & 1u);
            } else {
                cp_async_wait_pending<PD - 1>();
            }
            __syncthreads();
            if (!TMA) { issue(m + PD); cp_async_commit(); }
            if (!FINE && (k & 1)) {
                if (producer) stream_produce(items, smem, (int)lay.bars(), (m + 3) >> 1, pw_lane);
                continue;
            }
            if (FINE && m < STREAM_DF) continue;
            // window phase: the new row goes to slot N, the centre row (new one step ago) is C, the upper row is U
            const int N = PH3 ? (FINE ? ku % 3 : ku / 2) : 2, C = PH3 ? (N + 2) % 3 : 1, U = PH3 ? (N + 1) % 3 : 0;
            const int t = r0 - 3 + (FINE ? m - STREAM_DF : (m >> 1));     // newest row of this step
            if (XW && FINE) {
                lc.mb_rd = mbox + ((m + 1) & 1) * nch * PL_COUNT * 2 + wk * PL_COUNT * 2;
                lc.mb_wr = mbox + (m & 1) * nch * PL_COUNT * 2 + wk * PL_COUNT * 2;
            }
            // ---- row t from the staged ring (fine: the row itself; coarse: the 2x2 mean of fine rows 2t, 2t+1).
            //      clamp extension is applied HERE, when a row is produced: the row above row 0 becomes a copy of row 0
            //      (written into the centre slot the moment row 0 arrives), rows below the image repeat row H-1.
            if (t >= 0 && t < L.H) {
                if (FINE) {
                    z[N] = row_ld(zring + zs * Wp + lcol);
                } else {
                    const float* p0 = zring + zs * Wp + 2 * lcol;
                    const Row a0 = row_ld(p0), a1 = row_ld(p0 + 4), b0 = row_ld(p0 + Wp), b1 = row_ld(p0 + Wp + 4);
                    z[N].v[0] = 0.25f * (a0.v[0] + a0.v[1] + b0.v[0] + b0.v[1]);
                    z[N].v[1] = 0.25f * (a0.v[2] + a0.v[3] + b0.v[2] + b0.v[3]);
                    z[N].v[2] = 0.25f * (a1.v[0] + a1.v[1] + b1.v[0] + b1.v[1]);
                    z[N].v[3] = 0.25f * (a1.v[2] + a1.v[3] + b1.v[2] + b1.v[3]);
                }
                if (t == 0) z[C] = z[N];
            } else {
                z[N] = t < 0 ? row_zero() : z[C];
            }
            mb_post<XW && FINE>(z[N], lc, PL_Z);
            // ---- S at row t-1 (clamp-extended like z)
            {
                const int r = t - 1;
                if (r >= 0 && r < L.H) {
                    float l, rr;
                    nb_lr<false, XW && FINE>(z[C], l, rr, lc, PL_Z);
                    sB[N] = w_S(L.kT, z[C], z[U], z[N], l, rr);
                    if (GLR) sA[N] = w_S(L.kL, z[C], z[U], z[N], l, rr);
                    if (r == 0) { sB[C] = sB[N]; if (GLR) sA[C] = sA[N]; }
                } else if (r >= L.H) {
                    sB[N] = sB[C];
                    if (GLR) sA[N] = sA[C];
                }
                if (GLR) mb_post<XW && FINE>(sA[N], lc, PL_SA);
                mb_post<XW && FINE>(sB[N], lc, PL_SB);
            }
            // ---- L and the GTV cores at row t-2 (zero rows outside the image)
            {
                const int r = t - 2;
                if (r >= 0 && r < L.H) {
                    float l, rr;
                    if (GLR) {
                        nb_lr<false, XW && FINE>(sA[C], l, rr, lc, PL_SA);
                        Row w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) w[e] = row_ld(wrow(SM::PL_WL + e, r));
                        lA[N] = w_L(sA[C], sA[U], sA[N], l, rr, w);
                    }
                    nb_lr<false, XW && FINE>(sB[C], l, rr, lc, PL_SB);
                    const float* crow = wrow(SM::PL_C, r);
                    const Row cr = row_ld(crow), cd = row_ld(wrow(SM::PL_C + 1, r));
                    const float cr_left = lc.first ? 0.f : crow[-1];
                    oB[N] = w_core_lin(sB[C], sB[U], sB[N], l, rr, cr, cr_left, cd, cDp);
                    cDp = cd;
                    if (THR) {
                        // raw weights: plane U of this row was read one step ago (it is staged one row ahead), plane D of
                        // the row above is last step's own[3]
                        RawW rw;
                        rw.own[0] = wU_next;
                        rw.in[0] = wD_prev;
                        const float* pL = wrow(SM::PL_WT + 1, r);
                        const float* pR = wrow(SM::PL_WT + 2, r);
                        rw.own[1] = row_ld(pL);
                        rw.own[2] = row_ld(pR);
                        rw.own[3] = row_ld(wrow(SM::PL_WT + 3, r));
                        rw.in[3] = r + 1 < L.H ? row_ld(wrow(SM::PL_WT + 0, r + 1)) : row_zero();   // edge U of the lower neighbour
                        const float wr_m1 = lc.first ? 0.f : pR[-1];                       // edge R of the left neighbour
                        const float wl_p4 = lc.last ? 0.f : pL[4];                         // edge L of the right neighbour
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            rw.in[1].v[j] = j == 0 ? wr_m1 : rw.own[2].v[j - 1];
                            rw.in[2].v[j] = j == 3 ? wl_p4 : rw.own[1].v[j + 1];
                        }
                        oT[N] = w_core_thr(sB[C], sB[U], sB[N], l, rr, rw, L.Gam);
                        wU_next = rw.in[3];
                        wD_prev = rw.own[3];
                    }
                } else {
                    if (GLR) lA[N] = row_zero();
                    oB[N] = row_zero();
                    cDp = row_zero();
                    if (THR) {
                        oT[N] = row_zero();
                        wD_prev = row_zero();
                        wU_next = (r + 1 >= 0 && r + 1 < L.H) ? row_ld(wrow(SM::PL_WT + 0, r + 1)) : row_zero();
                    }
                }
                if (GLR) mb_post<XW && FINE>(lA[N], lc, PL_LA);
                mb_post<XW && FINE>(oB[N], lc, PL_OB);
                if (THR) mb_post<XW && FINE>(oT[N], lc, PL_OT);
            }
            // ---- St at row t-3 and the row's destination
            {
                const int r = t - 3;
                if (FINE ? (r >= R0 && r < R1) : (r >= 0 && r < L.H)) {
                    float l, rr;
                    nb_lr<true, XW && FINE>(oB[C], l, rr, lc, PL_OB);
                    Row Az = w_St(L.kT, oB[C], oB[U], oB[N], l, rr);
#pragma unroll
                    for (int j = 0; j < 4; ++j) Az.v[j] *= L.aT;
                    if (GLR) {
                        nb_lr<true, XW && FINE>(lA[C], l, rr, lc, PL_LA);
                        const Row gl_ = w_St(L.kL, lA[C], lA[U], lA[N], l, rr);
#pragma unroll
                        for (int j = 0; j < 4; ++j) Az.v[j] += L.aL * gl_.v[j];
                    }
                    Row rT = row_zero();
                    if (THR) {
                        nb_lr<true, XW && FINE>(oT[C], l, rr, lc, PL_OT);
                        rT = w_St(L.kT, oT[C], oT[U], oT[N], l, rr);
#pragma unroll
                        for (int j = 0; j < 4; ++j) rT.v[j] *= L.aT;
                    }
                    if (!FINE) {
                        // coarse: hand 0.25 * (coarse term) to the fine epilogue of rows 2r, 2r+1
                        if (lc.active) {
                            float* slot = cring + (r & 1) * NRING * Wpc + lcol;
                            float v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = 0.25f * Az.v[j];
                            st4(slot, v);
                            if (THR) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) v[j] = 0.25f * rT.v[j];
                                st4(slot + Wpc, v);
                            }
                        }
                    } else {
                        const float* slot = cring + ((r >> 1) & 1) * NRING * Wpc + (lcol >> 1);
                        GLR_CHECK_ALIGN(slot, 8);
                        const float2 cz = *reinterpret_cast<const float2*>(slot);
                        const Row zq = row_ld(zring + zqs * Wp + lcol);      // row t-3 is still in the ring
#pragma unroll
                        for (int j = 0; j < 4; ++j) Az.v[j] += zq.v[j] + (j < 2 ? cz.x : cz.y);
                        if (THR) {
                            const float2 ctv = *reinterpret_cast<const float2*>(slot + Wpc);
#pragma unroll
                            for (int j = 0; j < 4; ++j) rT.v[j] += (j < 2 ? ctv.x : ctv.y);
                        }
                        const float* ops = opring + (r & (STREAM_OPR - 1)) * Wp + lcol;
                        Row in0 = row_zero(), in1 = row_zero(), in2 = row_zero(), o0, o1, o2;
                        if (MODE == MODE_X2 || has_skip) in0 = row_ld(ops);
                        if (MODE == MODE_X3) { in1 = row_ld(ops + STREAM_OPR * Wp); in2 = row_ld(ops + 2 * STREAM_OPR * Wp); }
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
                        if (lc.active && lc.col0 >= ct.v0 && lc.col0 < ct.v1) {
                            const size_t gi = off + (size_t)r * W + lc.col0;
                            st4(a.out0 + gi, o0.v);
                            if (MODE == MODE_X2) { st4(a.out1 + gi, o1.v); st4(a.out2 + gi, o2.v); }
                        }
                    }
                }
            }
            // ---- advance the ring slots
            if (FINE) {
                zs = zs + 1 == ZR ? 0 : zs + 1;
                zqs = zqs + 1 == ZR ? 0 : zqs + 1;
            } else {
                zs = zs + 2 == ZR ? 0 : zs + 2;
            }
            if (!PH3) {
                z[U] = z[C]; z[C] = z[N]; sB[U] = sB[C]; sB[C] = sB[N]; oB[U] = oB[C]; oB[C] = oB[N];
                if (GLR) { sA[U] = sA[C]; sA[C] = sA[N]; lA[U] = lA[C]; lA[C] = lA[N]; }
                if (THR) { oT[U] = oT[C]; oT[C] = oT[N]; }
            }
        }
    }
    if (!TMA) cp_async_wait_all();
}

template <int MODE, bool XW, bool TMA>
__global__ void __launch_bounds__(STREAM_MAXT, STREAM_MINB) k_stream_fwd(StreamFwdArgs a) {
    GLR_SMEM_DECL(smem);
    const int W = a.s.W;
    const int GL = XW ? 64 : (W <= 32 ? 8 : W <= 64 ? 16 : 32), GLc = GL / 2;
    const int NTF = (a.nch * GL + 31) & ~31;      // fine threads first, then coarse threads (roles never share a warp)
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    // clear the staging memory once (rows outside the image are never copied, and what is read in their place must be
    // finite), initialise the mbarriers, and make both visible to the TMA unit
    {
        StreamSmem<MODE> lay; lay.Wp = 4 * GL; lay.nch = a.nch;
        const int n4 = (int)(lay.total() / 4);
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += NT) st4(smem + 4 * i, z4);
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < STREAM_NBAR; ++k) mbar_init(smem_advance(smem_addr(smem), (int)lay.bars() + 2 * k), 1);
        }
        mbar_fence_init();
        __syncthreads();
    }
    if (tid < NTF) stream_walk<MODE, XW, true, TMA>(a, smem, tid / GL, NTF / GL, tid % GL, GL);
    else stream_walk<MODE, XW, false, TMA>(a, smem, (tid - NTF) / GLc, (NT - NTF) / GLc, (tid - NTF) % GLc, GL);
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
unsigned long long g_glr_stream_launches = 0;   // streaming-path kernels launched (diagnostic)
extern "C" unsigned long long glrgtv_stream_launch_count(void) { return g_glr_stream_launches; }
int g_glr_block_path = 0;   // 0 auto, 1 plane kernels only, 2 streaming kernels (error when the shape is not eligible)
extern "C" int glrgtv_set_block_path(int mode) {
    if (mode < 0 || mode > 2) return GLRGTV_ERR_UNSUPPORTED;
    g_glr_block_path = mode;
    return GLRGTV_OK;
}

// planes wider than a 64-lane walker are cut into column strips of STREAM_STRIP valid columns (4K inference, wide patches)
int glr_stream_fwd_eligible(const glrgtv_shape* s) {
    return s->W % 8 == 0 && s->H % 2 == 0 && s->H >= 2;
}

struct StreamPlan {
    int GL, nch, threads, band_rows, n_bands, n_strips;
};
static StreamPlan stream_plan(const glrgtv_shape& s, int rows) {
    StreamPlan p;
    p.GL = s.W > 128 ? 64 : s.W > 64 ? 32 : s.W > 32 ? 16 : 8;
    p.n_strips = s.W > 256 ? (s.W + STREAM_STRIP - 1) / STREAM_STRIP : 1;
    p.nch = 1;
    for (int n = 1; n <= s.F; ++n) {
        if (s.F % n) continue;
        const int thr = ((n * p.GL + 31) & ~31) + ((n * p.GL / 2 + 31) & ~31);
        if (thr <= STREAM_MAXT) p.nch = n;
    }
    p.threads = ((p.nch * p.GL + 31) & ~31) + ((p.nch * p.GL / 2 + 31) & ~31);
    // whole-height bands unless the grid would leave SMs idle
    const long ctas = (long)s.B * s.G * (s.F / p.nch) * p.n_strips;
    int bands = 1;
    while (ctas * bands < 592 && rows / (bands * 2) >= 64) bands *= 2;
    p.band_rows = ((rows + bands - 1) / bands + 1) & ~1;
    p.n_bands = (rows + p.band_rows - 1) / p.band_rows;
    return p;
}

// 0: automatic (measured on B200: the TMA producer wins for the thresholded stage X2 - ten weight planes per level - on
// rows of 512 bytes and more; the per-thread cp.async loader wins elsewhere), 1: cp.async everywhere, 2: TMA everywhere
int g_glr_stream_loader = 0;
extern "C" int glrgtv_set_stream_loader(int mode) {
    if (mode < 0 || mode > 2) return GLRGTV_ERR_UNSUPPORTED;
    g_glr_stream_loader = mode;
    return GLRGTV_OK;
}

template <int MODE, bool XW, bool TMA>
static int launch_stream_kernel(const StreamFwdArgs& a, const StreamPlan& p, long blocks, void* stream) {
    StreamSmem<MODE> lay; lay.Wp = 4 * p.GL; lay.nch = p.nch;
    const size_t smem = lay.total() * sizeof(float);
    if (smem > 227 * 1024) return GLRGTV_ERR_UNSUPPORTED;
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_stream_fwd<MODE, XW, TMA>, smem, optin)) return rc_;
#endif
    ++g_glr_stream_launches;
    GLR_LAUNCH_FIBERS((k_stream_fwd<MODE, XW, TMA>), dim3((unsigned)blocks), p.threads, smem, stream, a);
    return GLRGTV_OK;
}
// second-generation pair walkers (fw2.cuh; instantiated in fw2_<stage>.cu)
bool glr_fw2_wanted(int mode, const glrgtv_shape* s);
bool glr_fw2_eligible(const glrgtv_shape* s);
extern template int glr_fw2_stage<FW_BA>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
extern template int glr_fw2_stage<FW_X1>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
extern template int glr_fw2_stage<FW_X2>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
extern template int glr_fw2_stage<FW_X3>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
static int glr_fw2_dispatch(int mode, const StreamFwdArgs& a, void* stream) {
    F2Args f = {};
    f.s = a.s; f.p = a.p;
    f.z = a.z; f.cT = a.cT0; f.wT = a.wT0; f.wL = a.wL0;
    f.out0 = a.out0; f.out1 = a.out1; f.out2 = a.out2;
    switch (mode) {
        case MODE_BA: return glr_fw2_stage<FW_BA>(f, a.cT1, a.wT1, a.wL1, a.vc, a.row_base, a.row_end, stream);
        case MODE_X1: return glr_fw2_stage<FW_X1>(f, a.cT1, a.wT1, a.wL1, a.vc, a.row_base, a.row_end, stream);
        case MODE_X2: f.op0 = a.y; return glr_fw2_stage<FW_X2>(f, a.cT1, a.wT1, a.wL1, a.vc, a.row_base, a.row_end, stream);
        default: f.op0 = a.bB_in; f.op1 = a.r1_in; f.op2 = a.p.skip ? a.y : nullptr;
                 return glr_fw2_stage<FW_X3>(f, a.cT1, a.wT1, a.wL1, a.vc, a.row_base, a.row_end, stream);
    }
}

template <int MODE>
static int launch_stream_stage(StreamFwdArgs a, void* stream) {
    const glrgtv_shape& s = a.s;
    if (a.vc != nullptr && glr_fw2_wanted(MODE, &s) && glr_fw2_eligible(&s)) {
        GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_BA + MODE, stream);
        const int rc2 = glr_fw2_dispatch(MODE, a, stream);
        GLR_PROF_END(GLRGTV_SLOT_FWD_BA + MODE, stream);
        return rc2;
    }
    const StreamPlan p = stream_plan(s, a.row_end - a.row_base);
    a.nch = p.nch; a.band_rows = p.band_rows; a.n_bands = p.n_bands; a.n_strips = p.n_strips;
    const long blocks = (long)s.B * s.G * (s.F / p.nch) * p.n_bands * p.n_strips;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_BA + MODE, stream);
    // the TMA producer keeps at most 32 * STREAM_MAXI row copies per batch
    const int nf = StreamSmem<MODE>::NPL + p.nch * StreamSmem<MODE>::NOP, nc = StreamSmem<MODE>::NPL + 2 * p.nch;
    const bool want_tma = g_glr_stream_loader == 2 || (g_glr_stream_loader == 0 && MODE == MODE_X2 && s.W >= 128);
    const bool tma = want_tma && 2 * nf + nc <= 32 * STREAM_MAXI && p.n_strips == 1;      // (the TMA producer copies whole rows)
    const int rc = p.GL == 64 ? (tma ? launch_stream_kernel<MODE, true, true>(a, p, blocks, stream) : launch_stream_kernel<MODE, true, false>(a, p, blocks, stream))
                              : (tma ? launch_stream_kernel<MODE, false, true>(a, p, blocks, stream) : launch_stream_kernel<MODE, false, false>(a, p, blocks, stream));
    GLR_PROF_END(GLRGTV_SLOT_FWD_BA + MODE, stream);
    return rc ? rc : GLR_CHECK_LAUNCH();
}

int glr_launch_gtv_coeffs(const glrgtv_shape& s, const float* w, float* c, void* stream) {
    const long planes = (long)s.B * s.G;
    const size_t total = (size_t)planes * s.H * s.W;
#ifdef GLRGTV_EMU
    const unsigned blocks = 1;
#else
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    GLR_LAUNCH(k_gtv_coeffs, dim3(blocks ? blocks : 1), 256, 0, stream, (int)planes, s.H, s.W, w, c);
    GLR_PROF_END(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    return GLR_CHECK_LAUNCH();
}

// the four solver stages on the streaming kernels; weights and coefficient planes are already in `sv`
int glr_stream_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x, float* out,
                         const glrgtv_block_saved* sv, void* stream) {
    int rc;
    StreamFwdArgs a;
    a.s = *s; a.p = *p;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1; a.cT0 = sv->cT0; a.cT1 = sv->cT1; a.vc = sv->vc;
    a.y = nullptr; a.bB_in = nullptr; a.r1_in = nullptr; a.out1 = nullptr; a.out2 = nullptr;
    a.nch = 1; a.band_rows = s->H; a.n_bands = 1; a.n_strips = 1; a.row_base = 0; a.row_end = s->H;
    a.z = x; a.out0 = sv->bA;
    if ((rc = launch_stream_stage<MODE_BA>(a, stream))) return rc;
    a.z = sv->bA; a.out0 = sv->x1;
    if ((rc = launch_stream_stage<MODE_X1>(a, stream))) return rc;
    a.z = sv->x1; a.y = x; a.out0 = sv->x2; a.out1 = sv->bB; a.out2 = sv->r1;
    if ((rc = launch_stream_stage<MODE_X2>(a, stream))) return rc;
    a.z = sv->x2; a.y = x; a.bB_in = sv->bB; a.r1_in = sv->r1; a.out0 = out; a.out1 = nullptr; a.out2 = nullptr;
    return launch_stream_stage<MODE_X3>(a, stream);
}

// One solver stage on the rows [row0, row1) of the plane (even bounds).  The rows outside the range must already hold
// valid data in the stage's inputs wherever the chain reaches them (8 rows on either side, see glrgtv.h): this is what a
// spatially sharded caller exchanges between stages.
int glr_stream_block_fwd_stage(int stage, const glrgtv_shape* s, const glrgtv_block_params* p, const float* x, float* out,
                               const glrgtv_block_saved* sv, int row0, int row1, void* stream) {
    StreamFwdArgs a;
    a.s = *s; a.p = *p;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1; a.cT0 = sv->cT0; a.cT1 = sv->cT1; a.vc = sv->vc;
    a.y = nullptr; a.bB_in = nullptr; a.r1_in = nullptr; a.out1 = nullptr; a.out2 = nullptr;
    a.nch = 1; a.band_rows = s->H; a.n_bands = 1; a.n_strips = 1; a.row_base = row0; a.row_end = row1;
    switch (stage) {
        case MODE_BA: a.z = x; a.out0 = sv->bA; return launch_stream_stage<MODE_BA>(a, stream);
        case MODE_X1: a.z = sv->bA; a.out0 = sv->x1; return launch_stream_stage<MODE_X1>(a, stream);
        case MODE_X2: a.z = sv->x1; a.y = x; a.out0 = sv->x2; a.out1 = sv->bB; a.out2 = sv->r1; return launch_stream_stage<MODE_X2>(a, stream);
        case MODE_X3: a.z = sv->x2; a.y = x; a.bB_in = sv->bB; a.r1_in = sv->r1; a.out0 = out; return launch_stream_stage<MODE_X3>(a, stream);
    }
    return GLRGTV_ERR_UNSUPPORTED;
}
