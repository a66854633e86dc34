// bw2.cuh - second-generation streaming backward stages of LocalLowpassFilteringBlock / MixtureGTVGLR:
// one walker body, packed fp32 (FFMA2) arithmetic, and the edge-weight gradients folded into the adjoint walk.
//
// What changed against block_stream_bwd.cu + block_gw.cu (round 1: 168 registers, four role bodies, ~400 instructions per
// element, a separate tiled pass per stage for the edge-weight gradients):
//   * ONE walker per channel does both operator families (GTV and GLR) of ONE resolution; the half-resolution branch is
//     its own launch of the same kernel (COARSE), run first, whose result Vc the full-resolution launch adds in its epilogue
//     (A^T g = g + F0^T g + P^T[F1^T (P g)], SURVEY B.8: P and P^T are exact mutual adjoints).  No role dispatch, no
//     coarse-to-fine ring, one instruction stream that fits the instruction cache.
//   * a lane owns a PAIR of adjacent pixels and every row lives in an aligned 64-bit register pair: stencil arithmetic is
//     fma.rn.f32x2 / add.f32x2 / mul.f32x2 (SASS FFMA2 / FADD2 / FMUL2, sm_100 only) with scalar taps broadcast for free;
//   * the forward chain is NOT recomputed past the cores: every parameter gradient that needed A z is rewritten with the
//     adjoint identity <g, A z> = <A^T g, z> (the walker produces A^T g anyway), so St o never has to be formed;
//   * the stage input and the upstream gradient are read from the shared-memory rings (cp.async, two steps ahead) at the
//     rows a stencil needs them - left / right neighbours included - instead of living in register windows;
//   * edge-weight gradients: every walker posts its per-edge products of the first-level stencils (s, h) to shared memory,
//     the CTA - which holds ALL channels of one (batch, graph) - sums them over the channels one step later and writes the
//     gradient planes with red.global.add (every stage of a level accumulates into one buffer, zeroed by the caller).
//
// Per step t the walker handles: rows t of z and g (staged), first-level stencils at row t-1, cores + products at row t-2,
// the scatter form of S' at row t-3 and the stage epilogue at row t-4.
//
//   MODE    chain input g        scale   output                                   reference lines (V1X0)
//   X3      gout                 -a2 s1  gx2 = s1 gout + c W ; gA ; gB            788-790, 985-988
//   X2A     gA                   1       gx1 = gx2 + W                            784-786 (r1 = bB - A x1)
//   X2B     gB                   1       gx1 += W_thr   (thresholded GTV only)     757-781
//   X1      gx1                  -a0     gbA = (1 + a0) gx1 + c W                 751-753
//   BA      gbA                  1       gx  = gbA + W + gB + s0 gout             738-749
// with W = A^T g (X2B, BA: only the GTV part, without the identity), gA = -(b2 a2 s1 gout + a1 gx2) and
// gB = (1 + b2) a2 s1 gout + a1 gx2 (the upstream gradients of r1 and bB).
#pragma once
#include "stream.cuh"
#include "stream_bwd.cuh"

// ------------------------------------------------------------------ packed pairs
#ifdef GLRGTV_EMU
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#else
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
#endif
__device__ __forceinline__ float2 pset(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 pfmas(float2 a, float s, float2 c) { return pfma(a, pset(s), c); }      // a * s + c
__device__ __forceinline__ float2 pmuls(float2 a, float s) { return pmul(a, pset(s)); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return pfma(b, pset(-1.f), a); }             // a - b
__device__ __forceinline__ float2 pld(const float* p) { GLR_CHECK_ALIGN(p, 8); return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void pst(float* p, float2 v) { GLR_CHECK_ALIGN(p, 8); *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ float2 pzero() { return make_float2(0.f, 0.f); }
// the pair shifted by one pixel: left neighbours (l, c.x), right neighbours (c.y, r)
__device__ __forceinline__ float2 pshl(float2 c, float l) { return make_float2(l, c.x); }
__device__ __forceinline__ float2 pshr(float2 c, float r) { return make_float2(c.y, r); }

#define B2_PD 2          // steps between issuing a copy and reading it
#define B2_ZR 8          // ring rows of the stage input / upstream gradient (rows t-4 .. t+2 are live)
#define B2_ZRC 4         // COARSE: ring rows of the staged full-resolution rows (pooled at arrival into rings of B2_ZR rows)
#define B2_WR 4          // ring rows of the weight planes
#define B2_OPR 4         // ring rows of the epilogue operands
#define B2_NFLD 6        // seam mailbox fields: s_T s_L h_T h_L gs_T gs_L
#ifndef B2_MAXT
#define B2_MAXT 384
#endif
#ifndef B2_MINB
#define B2_MINB 1
#endif

enum { B2_ST = 0, B2_SL = 1, B2_HT = 2, B2_HL = 3, B2_GT = 4, B2_GL = 5 };

// Shared-memory layout in floats.  Every row of every ring is RW = 2 L floats with L (lanes per row) a power of two, so a ring
// slot is a shift and a mask away from the step counter; L > 32: the walker of one channel spans several warps.
template <int MODE, bool COARSE>
struct B2Smem {
    static constexpr bool HAS_L = MODE == BW_X3 || MODE == BW_X2A || MODE == BW_X1, THR = MODE == BW_X2B;
    static constexpr int NPT = 4, NPLW = NPT + (HAS_L ? 4 : 0), NGW = 4 + (HAS_L ? 4 : 0);
    // epilogue operand tensors of the full-resolution launch (X3: r1, bB, x; X2A: r1, gx2; X2B: gx1; BA: gB, gout)
    static constexpr int NOP = COARSE ? 0 : MODE == BW_X3 ? 3 : (MODE == BW_X2A || MODE == BW_BA) ? 2 : MODE == BW_X2B ? 1 : 0;
    int NCH, L;       // channels per CTA, lanes per row
    __host__ __device__ int rw() const { return 2 * L; }
    __host__ __device__ size_t zsig() const { return 0; }                                                    // [NCH][ZR][RW] level rows of z
    __host__ __device__ size_t gsig() const { return zsig() + (size_t)NCH * B2_ZR * rw(); }                 // [NCH][ZR][RW] level rows of g
    __host__ __device__ size_t zstage() const { return gsig() + (size_t)NCH * B2_ZR * rw(); }               // COARSE [NCH][ZRC][2][2 RW] full-resolution rows
    __host__ __device__ size_t gstage() const { return zstage() + (COARSE ? (size_t)NCH * B2_ZRC * 4 * rw() : 0); }
    __host__ __device__ size_t opring() const { return gstage() + (COARSE ? (size_t)NCH * B2_ZRC * 4 * rw() : 0); }  // [NOP][NCH][4][RW]
    __host__ __device__ size_t vcring() const { return opring() + (size_t)NOP * NCH * B2_OPR * rw(); }              // fine: [NCH][4][RW/2]
    __host__ __device__ size_t wring() const { return vcring() + (COARSE ? 0 : (size_t)NCH * B2_OPR * (rw() / 2)); }  // [NPLW][WR][RW]
    __host__ __device__ size_t gwpost() const { return wring() + (size_t)NPLW * B2_WR * rw(); }             // [2][NCH][NGW][RW]
    __host__ __device__ size_t mbox() const { return gwpost() + (size_t)2 * NCH * NGW * rw(); }             // [2][NCH][warps per row][NFLD][2]
    __host__ __device__ size_t scratch() const { return (mbox() + (size_t)2 * NCH * ((L + 31) / 32) * B2_NFLD * 2 + 3) & ~(size_t)3; }
    __host__ __device__ size_t total() const { return scratch() + (size_t)NGW * rw(); }                     // scratch rows: posts of padding threads
    __host__ __device__ size_t bytes() const { return total() * sizeof(float); }
};

// 16-byte asynchronous copy global -> shared; !valid: the 16 bytes are ZERO-FILLED (src-size 0) and src is not read
#ifdef GLRGTV_EMU
__device__ __forceinline__ void b2_cp16(float* dst, const float* src, bool valid) {
    static const float zeros[4] = {0.f, 0.f, 0.f, 0.f};
    cp_async16(dst, valid ? src : zeros);
}
#else
__device__ __forceinline__ void b2_cp16(float* dst, const float* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 16 : 0));
}
#endif

// accumulate a pair into global memory without a return value (SASS RED.E.ADD.F32x2)
#ifdef GLRGTV_EMU
__device__ __forceinline__ void b2_red2(float* dst, float2 v) { dst[0] += v.x; dst[1] += v.y; }
#else
__device__ __forceinline__ void b2_red2(float* dst, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(v.x), "f"(v.y) : "memory");
}
#endif

struct B2Lane {
    int width;
    bool first, last, seamL, seamR, post0, post31;
};

// scalars left / right of the pair c from the lanes next door.  When the walker spans several warps the seam scalars come from
// the mailbox `mb` (this warp's slot, written one step ago: every horizontally exchanged row is at least one step old).
template <bool ZERO, bool XW>
__device__ __forceinline__ void b2_nb(const float2& c, float& l, float& r, const B2Lane& lc, const float* mb, int field) {
    const float up = __shfl_up_sync(0xffffffffu, c.y, 1, lc.width);
    const float dn = __shfl_down_sync(0xffffffffu, c.x, 1, lc.width);
    l = lc.first ? (ZERO ? 0.f : c.x) : up;
    r = lc.last ? (ZERO ? 0.f : c.y) : dn;
    if (XW) {
        if (lc.seamL) l = mb[2 * field + 1 - 2 * B2_NFLD];           // last scalar of the warp to the left
        if (lc.seamR) r = mb[2 * field + 2 * B2_NFLD];               // first scalar of the warp to the right
    }
}
template <bool XW>
__device__ __forceinline__ void b2_post(const float2& c, const B2Lane& lc, float* mb, int field) {
    if (XW) {
        if (lc.post0) mb[2 * field] = c.x;
        if (lc.post31) mb[2 * field + 1] = c.y;
    }
}

// S (V1X0:177-195) on pairs: k_c c + k_R right + k_D down + k_U up + k_L left; cl / cr = the pair shifted left / right
__device__ __forceinline__ float2 b2_S(const StatsTaps& k, float2 c, float2 u, float2 d, float2 cl, float2 cr) {
    float2 o = pmuls(c, k.kc);
    o = pfmas(cr, k.kr, o);
    o = pfmas(d, k.kd, o);
    o = pfmas(u, k.ku, o);
    o = pfmas(cl, k.kl, o);
    return o;
}

// LGT / NCHT: log2 of the shared-memory row length and the channels per CTA as COMPILE-TIME constants (every ring offset is then
// an immediate and the cross-channel sums unroll); 0 = taken from the arguments (any shape, slower).  XWG: the generic
// kernel's "walker spans several warps" flag.
template <int MODE, bool COARSE, int LGT, int NCHT, bool XWG>
__global__ void __launch_bounds__(B2_MAXT, B2_MINB) k_bw2(B2Args a) {
    GLR_SMEM_DECL(smem);
    using SM = B2Smem<MODE, COARSE>;
    constexpr bool HAS_L = SM::HAS_L, THR = SM::THR;
    constexpr bool XW = LGT ? (LGT >= 7) : XWG;
    constexpr int NPT = SM::NPT, NPLW = SM::NPLW, NGW = SM::NGW, PD = B2_PD;
    const int W = a.s.W, F = a.s.F, G = a.s.G;
    const int LH = COARSE ? a.s.H / 2 : a.s.H, LW = COARSE ? W / 2 : W;
    const int LG = LGT ? LGT : a.lg, RW = 1 << LG, L = RW >> 1, NCH = NCHT ? NCHT : a.nch;
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    const int ch = tid >> (LG - 1), lr = tid & (L - 1);
    const bool live = ch < NCH;
    const int chc = live ? ch : 0;
    int bid = (int)blockIdx.x;
    const int band = bid % a.n_bands; bid /= a.n_bands;
    const int part = bid % a.n_parts; bid /= a.n_parts;
    const int g = bid % G, b = bid / G;
    const int R0 = band * a.band_rows, R1 = R0 + a.band_rows < LH ? R0 + a.band_rows : LH;
    const int M = (R1 - R0) + 7;                   // steps: t = R0 - 3 .. R1 + 3, the epilogue trails by four rows
    const int c = g * F + part * NCH + chc;
    const size_t pl_f = (size_t)b * G * F + c;     // channel plane
    const size_t plane = (size_t)b * G + g;        // weight plane
    const int col0 = 2 * lr;
    const int NWR = (L + 31) >> 5;

    SM lay; lay.NCH = NCH; lay.L = L;
    float* const zsig = smem + lay.zsig() + ((size_t)chc * B2_ZR << LG) + col0;       // this lane's pair in slot 0
    float* const gsig = smem + lay.gsig() + ((size_t)chc * B2_ZR << LG) + col0;
    float* const wring = smem + lay.wring() + col0;
    const int M8 = (B2_ZR << LG) - 1, M4 = (B2_WR << LG) - 1;

    B2Lane lc;
    lc.width = L < 32 ? L : 32;
    const bool active = live && col0 < LW;
    lc.first = col0 == 0;
    lc.last = col0 + 2 >= LW;
    lc.seamL = XW && (lr & 31) == 0 && lr != 0;
    lc.seamR = XW && (lr & 31) == 31 && !lc.last;
    lc.post0 = XW && live && (lr & 31) == 0;
    lc.post31 = XW && live && (lr & 31) == 31;

    // ---- zero the shared memory once (inactive columns, the first ring rows)
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
    const float c23 = al2 * s1;
    const float ca = MODE == BW_X3 ? -c23 : MODE == BW_X1 ? -al0 : 1.f;       // scale of everything the chain produces
    const glrgtv_opparams& opT = COARSE ? a.p.gtv1 : a.p.gtv0;
    const glrgtv_opparams& opL = COARSE ? a.p.glr1 : a.p.glr0;
    const StatsTaps kT = glr_load_taps(opT.stats, c);
    const StatsTaps kL = HAS_L ? glr_load_taps(opL.stats, c) : kT;
    const float aT = expf(COARSE ? a.p.ro1[g] : a.p.ro0[g]);
    const float aL = HAS_L ? expf(COARSE ? a.p.mu1[g] : a.p.mu0[g]) : 0.f;
    const float Gam = THR ? expf(COARSE ? a.p.gamma1[g] : a.p.gamma0[g]) : 0.f;

    // ---- cp.async loader (every step issues its copies unconditionally: rows outside the image are CLAMPED for z and
    //      ZERO-FILLED for g and the weights, so the walker below needs no row tests).
    //      z / g, fine: the two lanes of a 16-byte piece share the copies (even lane: z, odd lane: g); coarse: every lane copies
    //      the two full-resolution rows of its own four columns.  Weights: up to two 16-byte pieces of plane rows per thread
    //      (4 NCH >= planes); plane U of a raw set runs one row ahead (its row r + 1 feeds the cores of row r).
    int tl = R0 - 3;                                  // the row the next issue() stages
    // (global addresses = tensor base + a 32-bit element offset that advances by one row per step)
    const float* zbase = a.z;
    unsigned zo = 0;                                  // offset of this thread's piece of row clamp(tl)
    float* zdst = smem;
    bool zok = false;
    if (COARSE) {
        zok = live && 4 * lr < W;
        zo = (unsigned)(pl_f * a.s.H * W) + (unsigned)(2 * glr_clampi(tl, 0, LH - 1) * W + (zok ? 4 * lr : 0));
        zdst = smem + lay.zstage() + ((size_t)chc * B2_ZRC * 4 << LG) + 4 * lr;
    } else {
        const int cc = 2 * (lr & ~1);
        zok = live && cc < W;
        zbase = (lr & 1) ? a.src : a.z;
        zo = (unsigned)(pl_f * a.s.H * W) + (unsigned)(glr_clampi(tl, 0, LH - 1) * W + (zok ? cc : 0));
        zdst = smem + ((lr & 1) ? lay.gsig() : lay.zsig()) + ((size_t)chc * B2_ZR << LG) + cc;
    }
    const bool zisz = COARSE || !(lr & 1);            // this thread's fine-level copy is a z row (clamped) / a g row (zero-filled outside)
    // epilogue operands (full resolution only): the even lane of a piece copies operand 0 and 2, the odd lane operand 1, for the
    // epilogue row of the step the z / g rows are staged for (tl - 4), through 32-bit element offsets from the tensor bases
    constexpr int NOP = SM::NOP;
    float* const opdst = smem + lay.opring() + ((size_t)chc * B2_OPR << LG) + 2 * (lr & ~1);
    unsigned ooff = (unsigned)(pl_f * a.s.H * W) + (unsigned)((tl - 4) * W + 2 * (lr & ~1));
    // the half-resolution launch's result: one 16-byte piece (4 of its pixels = the pairs of 4 lanes) per even epilogue row,
    // copied by the lane with lr % 4 == 3
    float* const vcdst = smem + lay.vcring() + ((size_t)chc * B2_OPR << (LG - 1)) + (lr & ~3);
    unsigned vcoff = (unsigned)(pl_f * (LH / 2) * (LW / 2)) + (unsigned)(((tl - 4) >> 1) * (LW / 2) + (lr & ~3));
    const float* wbase[2] = {a.wT, a.wT};
    unsigned wo[2] = {0, 0};
    float* wdst[2] = {smem, smem};
    bool wok[2] = {false, false};
    int wlead[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {                      // up to two pieces per thread (4 NCH >= planes)
        const int wi = tid + j * NT;
        const int pl = wi >> (LG - 2), piece = wi & ((L >> 1) - 1);
        if (pl < NPLW && 4 * piece < LW) {
            const bool isL = pl >= NPT;
            const int e = isL ? pl - NPT : pl;
            wlead[j] = e == 0 ? 1 : 0;
            wbase[j] = isL ? a.wL : a.wT;
            wo[j] = (unsigned)((plane * 4 + e) * LH * LW) + (unsigned)(glr_clampi(tl - 2 + wlead[j], 0, LH - 1) * LW + 4 * piece);
            wdst[j] = smem + lay.wring() + ((size_t)pl * B2_WR << LG) + 4 * piece;
            wok[j] = true;
        }
    }
    const int sdelta = COARSE ? (int)(lay.gstage() - lay.zstage()) : 0;
    auto issue = [&]() {                             // the copies of row tl; pointers advance to the next row unless it is clamped
        const bool in = (unsigned)tl < (unsigned)LH;
        if (zok) {
            if (COARSE) {
                float* d = zdst + ((tl & (B2_ZRC - 1)) * 4 << LG);
                cp_async16(d, a.z + zo); cp_async16(d + RW * 2, a.z + zo + W);
                b2_cp16(d + sdelta, a.src + zo, in); b2_cp16(d + sdelta + RW * 2, a.src + zo + W, in);
            } else {
                b2_cp16(zdst + ((tl & (B2_ZR - 1)) << LG), zbase + zo, in || zisz);
            }
        }
        if (tl >= 0 && tl < LH - 1) zo += COARSE ? 2 * W : W;
        if (NOP > 0 && zok && tl - 4 >= R0 && tl - 4 < R1) {
            float* d = opdst + (((tl - 4) & (B2_OPR - 1)) << LG);
            if (lr & 1) {
                if (NOP > 1) cp_async16(d + (NCH * B2_OPR << LG), a.op1 + ooff);
            } else {
                cp_async16(d, a.op0 + ooff);
                if (NOP > 2 && a.op2 != nullptr) cp_async16(d + (2 * NCH * B2_OPR << LG), a.op2 + ooff);
            }
        }
        ooff += W;
        if (!COARSE && live && (lr & 3) == 3 && (lr & ~3) < LW / 2 && !((tl - 4) & 1) && tl - 4 >= R0 - 1 && tl - 4 < R1)      // (a band may start on an odd row)
            cp_async16(vcdst + ((((tl - 4) >> 1) & (B2_OPR - 1)) << (LG - 1)), a.vc_in + vcoff);
        if ((tl - 4) & 1) vcoff += LW / 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int rw_ = tl - 2 + wlead[j];
            if (wok[j]) b2_cp16(wdst[j] + ((rw_ & (B2_WR - 1)) << LG), wbase[j] + wo[j], (unsigned)rw_ < (unsigned)LH);
            if (rw_ >= 0 && rw_ < LH - 1) wo[j] += LW;
        }
        ++tl;
    };

    // ---- walker state
    float2 sT[3], hT[3], sL[3], hL[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sT[k] = hT[k] = sL[k] = hL[k] = pzero();
    float2 gsT = pzero(), gsL = pzero();              // core outputs of the previous step (row t-3 at step t)
    float2 Vpend = pzero(), Vnext = pzero();          // scatter form of S': rows t-4 (+ pending k_U term) and t-3
    float2 wDpT = pzero(), wDpL = pzero();          // edge D of the row above (carried from the previous step)
    float2 tT[4], tL[4];                              // tap sums: c, R, D, U + L   
#pragma unroll
    for (int k = 0; k < 4; ++k) tT[k] = tL[k] = pzero();
    float2 sumT = pzero(), sumL = pzero(), gam = pzero();
    float2 fs[4];                                     // epilogue sums (X3: C1, A1, A4, D; X2A / X1: fs[0])
#pragma unroll
    for (int k = 0; k < 4; ++k) fs[k] = pzero();
    constexpr int N = 2, C = 1, U = 0;

    // per-step offsets, kept incrementally (no multiplications in the loop)
    int t = R0 - 3;
    int zt = (t & (B2_ZR - 1)) << LG;                 // ring slot of row t in the level rings
    int wt = ((t - 2) & (B2_WR - 1)) << LG;           // ring slot of row t-2 in the weight rings
    const int POSTS = (NCH * NGW) << LG, MBS = NCH * NWR * B2_NFLD * 2;
    // double buffers: what a step posts (gradient products, seam scalars) is read in the next step
    float* const postw0 = live ? smem + lay.gwpost() + ((size_t)chc * NGW << LG) + col0 : smem + lay.scratch() + col0;
    const float* const postr0 = smem + lay.gwpost() + col0;
    float* const mb0 = smem + lay.mbox() + (chc * NWR + (lr >> 5)) * B2_NFLD * 2;
    int pofs = 0, mofs = 0;                           // this step's buffer (0 / POSTS, 0 / MBS)
    // global element offset of (this channel, row t-4, col0) in the full-resolution tensors / the level tensors
    // (32-bit element offsets: tensors of up to 2^32 elements; they wrap harmlessly while the row is outside the band)
    unsigned goff = (unsigned)(pl_f * a.s.H * W) + (unsigned)((t - 4) * W + col0);
    unsigned loff = (unsigned)(pl_f * LH * LW) + (unsigned)((t - 4) * LW + col0);
    unsigned gwoff = (unsigned)(plane * 4 * LH * LW) + (unsigned)((t - 3) * LW + col0);
    const unsigned HWl = (unsigned)(LH * LW);

#pragma unroll
    for (int k = 0; k < PD; ++k) { issue(); cp_async_commit(); }

#pragma unroll 1
    for (int m = 0; m < M; ++m) {
        cp_async_wait_pending<PD - 1>();
        __syncthreads();
        issue();
        cp_async_commit();
        const int o1 = (zt - RW) & M8, o2 = (zt - 2 * RW) & M8, o3 = (zt - 3 * RW) & M8, o4 = (zt - 4 * RW) & M8;
        float* const postw = postw0 + (live ? pofs : 0);
        const float* const postr = postr0 + (POSTS - pofs);
        float* const mbw = mb0 + mofs;
        const float* const mbr = mb0 + (MBS - mofs);

        const int rf = t - 4;
        const bool fin = active && rf >= R0 && rf < R1;

        // ---- edge-weight gradients of row t-3: sum last step's posts over the channels (thread (ch, lr): planes ch, ch+NCH, ..)
        if (t - 3 >= R0 && t - 3 < R1 && active) {
#pragma unroll
            for (int k = 0; k < (NCHT ? (NGW + NCHT - 1) / NCHT : NGW); ++k) {
                const int pl = ch + k * NCH;
                if (pl < NGW) {
                    const float* pp = postr + (pl << LG);
                    float2 acc = pld(pp);
#pragma unroll
                    for (int f = 1; f < (NCHT ? NCHT : 1); ++f) acc = padd(acc, pld(pp + ((f * NGW) << LG)));
                    if (!NCHT) for (int f = 1; f < NCH; ++f) acc = padd(acc, pld(pp + ((f * NGW) << LG)));
                    acc = pmuls(acc, (!THR && pl < 4) ? 2.f * ca : ca);      // linear GTV stages post w D d: the gradient is twice that
                    b2_red2((pl < 4 ? a.gwT + pl * HWl : a.gwL + (pl - 4) * HWl) + gwoff, acc);
                }
            }
        }

        // ---- COARSE: pool the two staged full-resolution rows of row t into the level rings
        if (COARSE && live) {
            const float* zs = zdst + ((t & (B2_ZRC - 1)) * 4 << LG);
            float za[4], zb[4], ga[4], gb[4];
            ld4(zs, za); ld4(zs + 2 * RW, zb);
            ld4(zs + sdelta, ga); ld4(zs + sdelta + 2 * RW, gb);
            pst(zsig + zt, make_float2(0.25f * (za[0] + za[1] + zb[0] + zb[1]), 0.25f * (za[2] + za[3] + zb[2] + zb[3])));
            pst(gsig + zt, make_float2(0.25f * (ga[0] + ga[1] + gb[0] + gb[1]), 0.25f * (ga[2] + ga[3] + gb[2] + gb[3])));
        }

        // ---- rows t-1 and t-2 of z (clamp-extended) and g (zero-extended) with their left / right shifted pairs
        const float2 z0 = pld(zsig + zt), z1 = pld(zsig + o1), z2 = pld(zsig + o2), z3 = pld(zsig + o3);
        const float2 g0 = pld(gsig + zt), g1 = pld(gsig + o1), g2 = pld(gsig + o2), g3 = pld(gsig + o3);
        const float2 z1l = pshl(z1, lc.first ? z1.x : zsig[o1 - 1]), z1r = pshr(z1, lc.last ? z1.y : zsig[o1 + 2]);
        const float2 g1l = pshl(g1, lc.first ? 0.f : gsig[o1 - 1]), g1r = pshr(g1, lc.last ? 0.f : gsig[o1 + 2]);
        const float2 z2l = pshl(z2, lc.first ? z2.x : zsig[o2 - 1]), z2r = pshr(z2, lc.last ? z2.y : zsig[o2 + 2]);
        const float2 g2l = pshl(g2, lc.first ? 0.f : gsig[o2 - 1]), g2r = pshr(g2, lc.last ? 0.f : gsig[o2 + 2]);

        // ---- first-level stencils at row t-1: s = S z (clamp-extended), h = a S0 g (T: clamp-extended, L: zero-extended)
        {
            sT[N] = b2_S(kT, z1, z2, z0, z1l, z1r);
            hT[N] = pmuls(b2_S(kT, g1, g2, g0, g1l, g1r), aT);
            if (HAS_L) {
                sL[N] = b2_S(kL, z1, z2, z0, z1l, z1r);
                hL[N] = pmuls(b2_S(kL, g1, g2, g0, g1l, g1r), aL);
            }
            const int r = t - 1;
            if (r <= 0 || r >= LH) {               // image borders (warp-uniform, rare): s and T's h are clamp-extended, L's h zero-extended
                if (r < 0) { hT[N] = pzero(); hL[N] = pzero(); }
                else if (r == 0) { sT[C] = sT[N]; hT[C] = hT[N]; sL[C] = sL[N]; }
                else { sT[N] = sT[C]; hT[N] = hT[C]; sL[N] = sL[C]; hL[N] = pzero(); }
            }
            b2_post<XW>(sT[N], lc, mbw, B2_ST);
            b2_post<XW>(hT[N], lc, mbw, B2_HT);
            if (HAS_L) { b2_post<XW>(sL[N], lc, mbw, B2_SL); b2_post<XW>(hL[N], lc, mbw, B2_HL); }
        }

        // ---- S' in scatter form, fed with the core outputs of row t-3 (computed one step ago): finishes row t-4
        float2 Vdone;
        {
            const int r = t - 3;
            float l, rr;
            b2_nb<true, XW>(gsT, l, rr, lc, mbr, B2_GT);
            Vdone = pfmas(gsT, kT.ku, Vpend);
            float2 np = pfmas(gsT, kT.kc, Vnext);
            np = pfmas(pshl(gsT, l), kT.kr, np);
            np = pfmas(pshr(gsT, rr), kT.kl, np);
            float2 nn = pmuls(gsT, kT.kd);
            // the clamp-extended S collects its out-of-range taps on the border pixels
            np = pfma(gsT, make_float2(lc.first ? kT.kl : 0.f, lc.last ? kT.kr : 0.f), np);
            if (HAS_L) {
                b2_nb<true, XW>(gsL, l, rr, lc, mbr, B2_GL);
                Vdone = pfmas(gsL, kL.ku, Vdone);
                np = pfmas(gsL, kL.kc, np);
                np = pfmas(pshl(gsL, l), kL.kr, np);
                np = pfmas(pshr(gsL, rr), kL.kl, np);
                nn = pfmas(gsL, kL.kd, nn);
                np = pfma(gsL, make_float2(lc.first ? kL.kl : 0.f, lc.last ? kL.kr : 0.f), np);
            }
            if (r == 0 || r == LH - 1) {
                const float eT = (r == 0 ? kT.ku : 0.f) + (r == LH - 1 ? kT.kd : 0.f);
                np = pfmas(gsT, eT, np);
                if (HAS_L) { const float eL = (r == 0 ? kL.ku : 0.f) + (r == LH - 1 ? kL.kd : 0.f); np = pfmas(gsL, eL, np); }
            }
            Vpend = np;
            Vnext = nn;
        }

        // ---- cores at row t-2: o = core(s), gs = core'(h); tap sums; mu / ro; the edge-weight products
        {
            const int r = t - 2;
            const bool inimg = r >= 0 && r < LH;
            const bool cnt = active && r >= R0 && r < R1;
            const float2 zUL = padd(z3, z2l), gUL = padd(g3, g2l);
            float sl, sr, hl, hr;
            // ---------------- GTV
            b2_nb<false, XW>(sT[C], sl, sr, lc, mbr, B2_ST);
            b2_nb<false, XW>(hT[C], hl, hr, lc, mbr, B2_HT);
            const float2 dU = psub(sT[C], sT[U]), dL = psub(sT[C], pshl(sT[C], sl)), dR = psub(sT[C], pshr(sT[C], sr)), dD = psub(sT[C], sT[N]);
            const float2 DU = psub(hT[C], hT[U]), DL = psub(hT[C], pshl(hT[C], hl)), DR = psub(hT[C], pshr(hT[C], hr)), DD = psub(hT[C], hT[N]);
            float2 o, gs;
            {
                // raw weights around the pair: own[e] = w_e[q]; in[e] = the weight of the edge pointing from neighbour e back at q
                // (zero outside the image: the staged rows above / below it are zero-filled)
                const float* w0 = wring + wt;
                const float* pL = w0 + (1 * B2_WR << LG);
                const float* pR = w0 + (2 * B2_WR << LG);
                float2 own[4], in[4];
                own[0] = pld(w0); own[1] = pld(pL); own[2] = pld(pR); own[3] = pld(w0 + (3 * B2_WR << LG));
                in[0] = wDpT;
                in[3] = pld(wring + ((wt + RW) & M4));
                in[1] = make_float2(lc.first ? 0.f : pR[-1], own[2].x);
                in[2] = make_float2(own[1].y, lc.last ? 0.f : pL[2]);
                wDpT = own[3];
                const float2 dd[4] = {dU, dL, dR, dD}, DDv[4] = {DU, DL, DR, DD};
                if (!THR) {
                    // linear core Ct C with the symmetric coefficients c_e = w_e[q]^2 + w_opposite[neighbour e]^2 (self-adjoint)
                    o = pzero(); gs = pzero();
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 ce = pfma(own[e], own[e], pmul(in[e], in[e]));
                        o = pfma(ce, dd[e], o);
                        gs = pfma(ce, DDv[e], gs);
                        pst(postw + e * RW, pmul(own[e], pmul(DDv[e], dd[e])));
                    }
                } else {
                    o = pzero(); gs = pzero();
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float wa[2] = {own[e].x, own[e].y}, wb[2] = {in[e].x, in[e].y};
                        const float d2[2] = {dd[e].x, dd[e].y}, D2[2] = {DDv[e].x, DDv[e].y};
                        float ov[2], gv[2], pv[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float ta = wa[j] * d2[j], tb = wb[j] * d2[j];
                            const float pa = glr_phi(ta, Gam), da = glr_dphi(ta, Gam);
                            ov[j] = wa[j] * pa + wb[j] * glr_phi(tb, Gam);
                            gv[j] = D2[j] * (wa[j] * wa[j] * da + wb[j] * wb[j] * glr_dphi(tb, Gam));
                            pv[j] = D2[j] * pa + D2[j] * wa[j] * da * d2[j];
                            if (cnt && fabsf(ta) > Gam) { const float gm = D2[j] * wa[j] * (ta > 0.f ? -2.f : 2.f); if (j == 0) gam.x += gm; else gam.y += gm; }
                        }
                        o = padd(o, make_float2(ov[0], ov[1]));
                        gs = padd(gs, make_float2(gv[0], gv[1]));
                        pst(postw + e * RW, make_float2(pv[0], pv[1]));
                    }
                }
            }
            if (cnt) {
                tT[0] = pfma(gs, z2, tT[0]); tT[1] = pfma(gs, z2r, tT[1]); tT[2] = pfma(gs, z1, tT[2]); tT[3] = pfma(gs, zUL, tT[3]);
                const float2 oa = pmuls(o, aT);
                tT[0] = pfma(oa, g2, tT[0]); tT[1] = pfma(oa, g2r, tT[1]); tT[2] = pfma(oa, g1, tT[2]); tT[3] = pfma(oa, gUL, tT[3]);
                sumT = pfma(o, hT[C], sumT);
            }
            gsT = inimg ? gs : pzero();
            // ---------------- GLR
            if (HAS_L) {
                b2_nb<false, XW>(sL[C], sl, sr, lc, mbr, B2_SL);
                b2_nb<true, XW>(hL[C], hl, hr, lc, mbr, B2_HL);
                const float* w0 = wring + ((size_t)NPT * B2_WR << LG) + wt;
                const float* pL = w0 + (1 * B2_WR << LG);
                const float* pR = w0 + (2 * B2_WR << LG);
                const float2 wU = pld(w0), wLe = pld(pL), wRi = pld(pR), wD = pld(w0 + (3 * B2_WR << LG));
                const float2 inU = wDpL;                                                        // edge D of the upper neighbour
                const float2 inD = pld(wring + ((size_t)NPT * B2_WR << LG) + ((wt + RW) & M4));   // edge U of the lower neighbour (zero below the image)
                const float2 inL = make_float2(lc.first ? 0.f : pR[-1], wRi.x);                 // edge R of the left neighbour
                const float2 inR = make_float2(wLe.y, lc.last ? 0.f : pL[2]);                   // edge L of the right neighbour
                wDpL = wD;
                const float2 sLl = pshl(sL[C], sl), sLr = pshr(sL[C], sr);
                float2 acc = pmul(wU, sL[U]);
                acc = pfma(wLe, sLl, acc); acc = pfma(wRi, sLr, acc); acc = pfma(wD, sL[N], acc);
                const float2 oL = psub(sL[C], acc);
                // VJP of L wrt its input (h zero-extended): the out-of-range neighbours of a border pixel are the pixel itself
                float2 ad = pmul(inU, hL[U]);
                ad = pfma(inL, pshl(hL[C], hl), ad); ad = pfma(inR, pshr(hL[C], hr), ad); ad = pfma(inD, hL[N], ad);
                float2 self = make_float2(lc.first ? wLe.x : 0.f, lc.last ? wRi.y : 0.f);
                if (r == 0) self = padd(self, wU);
                if (r == LH - 1) self = padd(self, wD);
                float2 gL = psub(hL[C], ad);
                gL = psub(gL, pmul(self, hL[C]));
                const float2 nh = pmuls(hL[C], -1.f);
                pst(postw + 4 * RW, pmul(nh, sL[U])); pst(postw + 5 * RW, pmul(nh, sLl));
                pst(postw + 6 * RW, pmul(nh, sLr)); pst(postw + 7 * RW, pmul(nh, sL[N]));
                if (cnt) {
                    tL[0] = pfma(gL, z2, tL[0]); tL[1] = pfma(gL, z2r, tL[1]); tL[2] = pfma(gL, z1, tL[2]); tL[3] = pfma(gL, zUL, tL[3]);
                    const float2 oa = pmuls(oL, aL);
                    tL[0] = pfma(oa, g2, tL[0]); tL[1] = pfma(oa, g2r, tL[1]); tL[2] = pfma(oa, g1, tL[2]); tL[3] = pfma(oa, gUL, tL[3]);
                    sumL = pfma(oL, hL[C], sumL);
                }
                gsL = inimg ? gL : pzero();
            }
            b2_post<XW>(gsT, lc, mbw, B2_GT);
            if (HAS_L) b2_post<XW>(gsL, lc, mbw, B2_GL);
        }

        // ---- epilogue of row t-4
        if (fin) {
            if (COARSE) {
                pst(a.vc_out + loff, Vdone);
            } else {
                const float2 gq = pld(gsig + o4), zq = pld(zsig + o4);
                const float* ops = opdst - 2 * (lr & ~1) + col0 + ((rf & (B2_OPR - 1)) << LG);
                float2 q0 = pzero(), q1 = pzero(), q2 = pzero();
                if (NOP > 0) q0 = pld(ops);
                if (NOP > 1) q1 = pld(ops + (NCH * B2_OPR << LG));
                if (NOP > 2 && has_skip) q2 = pld(ops + (2 * NCH * B2_OPR << LG));
                const float vcv = vcdst[(((rf >> 1) & (B2_OPR - 1)) << (LG - 1)) + (lr & 3)];
                float2 Wp = pfmas(pset(vcv), 0.25f, Vdone);
                if (HAS_L) Wp = padd(Wp, gq);
                if (MODE == BW_X3) {              // gq gout, q0 r1, q1 bB, q2 x, zq x2
                    const float2 gx2 = pfmas(Wp, ca, pmuls(gq, s1));
                    pst(a.out0 + goff, gx2);
                    const float2 gB = pfmas(gx2, al1, pmuls(gq, c23 + be2 * c23));
                    pst(a.out1 + goff, psub(pmuls(gq, c23), gB));        // gA = c23 gout - gB
                    pst(a.out2 + goff, gB);
                    const float2 bb = pfmas(q0, be2, q1);
                    fs[0] = pfma(gq, bb, fs[0]); fs[0] = pfma(pmuls(Wp, -1.f), zq, fs[0]);
                    fs[1] = pfma(gq, q0, fs[1]);
                    fs[2] = pfma(gq, q2, fs[2]);
                    fs[3] = pfma(gq, zq, fs[3]);
                } else if (MODE == BW_X2A) {      // q0 r1, q1 gx2
                    pst(a.out0 + goff, padd(q1, Wp));
                    fs[0] = pfma(q1, q0, fs[0]);
                } else if (MODE == BW_X2B) {      // q0: the A part already in gx1
                    pst(a.out0 + goff, padd(q0, Wp));
                } else if (MODE == BW_X1) {       // gq gx1, zq bA
                    pst(a.out0 + goff, pfmas(Wp, ca, pmuls(gq, 1.f + al0)));
                    fs[0] = pfma(psub(gq, Wp), zq, fs[0]);
                } else {                          // BA: gq gbA, q0 gB, q1 gout
                    pst(a.out0 + goff, pfmas(q1, s0, padd(padd(gq, Wp), q0)));
                }
            }
        }
        // ---- rotate the windows, advance the offsets
        sT[U] = sT[C]; sT[C] = sT[N]; hT[U] = hT[C]; hT[C] = hT[N];
        if (HAS_L) { sL[U] = sL[C]; sL[C] = sL[N]; hL[U] = hL[C]; hL[C] = hL[N]; }
        ++t;
        zt = (zt + RW) & M8;
        wt = (wt + RW) & M4;
        goff += W; loff += LW; gwoff += LW;
        pofs = POSTS - pofs;
        mofs = MBS - mofs;
    }
    cp_async_wait_all();

    // ---- parameter gradients of this walker (scaled by the chain's factor): one reduction over the walker's lanes, one atomic each
    {
        auto red = [&](float2 v2, float* dst) {
            float v = live ? (v2.x + v2.y) * ca : 0.f;
            for (int o2 = lc.width / 2; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
            if ((lr & (lc.width - 1)) == 0 && live && v != 0.f) atomicAdd(dst, v);
        };
        const int Cn = G * F;
        auto taps = [&](const float2 (&tt)[4], float* dst) {
            // tap sums c, R, D, U+L -> p01, p02a, p02b, p03
            red(tt[0], dst + 0 * Cn + c);
            red(psub(tt[1], tt[0]), dst + 1 * Cn + c);
            red(psub(tt[2], tt[0]), dst + 2 * Cn + c);
            red(psub(pmuls(tt[0], 4.f), padd(padd(tt[1], tt[2]), tt[3])), dst + 3 * Cn + c);
        };
        taps(tT, COARSE ? a.gr.gtv1_stats : a.gr.gtv0_stats);
        red(sumT, (COARSE ? a.gr.ro1 : a.gr.ro0) + g);
        if (HAS_L) {
            taps(tL, COARSE ? a.gr.glr1_stats : a.gr.glr0_stats);
            red(sumL, (COARSE ? a.gr.mu1 : a.gr.mu0) + g);
        }
        if (THR) red(pmuls(gam, Gam), (COARSE ? a.gr.gamma1 : a.gr.gamma0) + g);
        if (!COARSE) {
            auto red1 = [&](float v, float* dst) {
                if (!live) v = 0.f;
                for (int o2 = lc.width / 2; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
                if ((lr & (lc.width - 1)) == 0 && live && v != 0.f) atomicAdd(dst, v);
            };
            if (MODE == BW_X3) {
                const float C1 = fs[0].x + fs[0].y, A1 = fs[1].x + fs[1].y, A4 = fs[2].x + fs[2].y, D = fs[3].x + fs[3].y;
                red1(s1 * C1, a.gr.alpha + 2 * G + g);
                red1(c23 * A1, a.gr.beta + 2 * G + g);
                if (has_skip && a.gr.skip) { red1(A4, a.gr.skip); red1(D + al2 * C1, a.gr.skip + 1); }
            } else if (MODE == BW_X2A) {
                red1(fs[0].x + fs[0].y, a.gr.alpha + G + g);
            } else if (MODE == BW_X1) {
                red1(fs[0].x + fs[0].y, a.gr.alpha + g);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side (templates: instantiated per MODE in bw2_*.cu so that the stage kernels compile in parallel)
// ---------------------------------------------------------------------------------------------------
extern unsigned long long g_glr_stream_launches;

static inline int b2_lanes(int lw) {                // lanes per row: pairs, padded to a power of two
    int p = 4;
    while (2 * p < lw) p *= 2;
    return p;
}
// channels per CTA: all of a graph's channels when they fit B2_MAXT threads, else the largest divisor of F that does (two
// CTAs then share a graph; their edge-weight gradient sums meet in global memory through red.global.add)
static inline int b2_nch(int F, int L) {
    for (int n = F; n >= 1; --n)
        if (F % n == 0 && n * L <= B2_MAXT) return n;
    return 0;
}
template <int MODE, bool COARSE>
static bool b2_fits(const glrgtv_shape* s) {
    const int L = b2_lanes(COARSE ? s->W / 2 : s->W), nch = b2_nch(s->F, L);
    if (nch < 1 || 4 * nch < B2Smem<MODE, COARSE>::NPLW || s->F / nch > 2) return false;
    B2Smem<MODE, COARSE> lay; lay.NCH = nch; lay.L = L;
    return lay.bytes() + 256 <= 227 * 1024;
}

template <int MODE, bool COARSE, int LGT, int NCHT, bool XWG>
static int b2_go(const B2Args& a, long blocks, int threads, size_t smem, void* stream) {
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_bw2<MODE, COARSE, LGT, NCHT, XWG>, smem, optin)) return rc_;
#endif
    GLR_LAUNCH_FIBERS((k_bw2<MODE, COARSE, LGT, NCHT, XWG>), dim3((unsigned)blocks), threads, smem, stream, a);
    return GLRGTV_OK;
}

template <int MODE, bool COARSE>
static int b2_launch(B2Args a, int slot, void* stream) {
    const glrgtv_shape& s = a.s;
    const int LH = COARSE ? s.H / 2 : s.H, LW = COARSE ? s.W / 2 : s.W;
    const int L = b2_lanes(LW);
    int lg = 0;
    while ((1 << lg) < 2 * L) ++lg;
    a.lg = lg;
    a.nch = b2_nch(s.F, L);
    if (a.nch < 1) return GLRGTV_ERR_UNSUPPORTED;
    a.n_parts = s.F / a.nch;
    const int threads = (a.nch * L + 31) & ~31;
    B2Smem<MODE, COARSE> lay; lay.NCH = a.nch; lay.L = L;
    const size_t smem = lay.bytes();
    if (threads > B2_MAXT || smem > 227 * 1024) return GLRGTV_ERR_UNSUPPORTED;
    // row bands: the split that minimises (waves of CTAs) x (steps of one CTA); a band pays 7 pipeline-fill steps
    int occ = (int)(227 * 1024 / (smem + 1024));
    if (occ > B2_MAXT / threads) occ = B2_MAXT / threads;       // register file: the kernels are built for B2_MAXT threads per SM
    if (occ < 1) occ = 1;
    int bands = 1;
    {
        const long base = (long)s.B * s.G * a.n_parts, slots = 148L * occ;
        long best = -1;
        for (int bnd = 1; bnd <= 8 && LH / bnd >= 24; bnd *= 2) {
            const long cost = ((base * bnd + slots - 1) / slots) * ((LH + bnd - 1) / bnd + 7);
            if (best < 0 || cost < best) { best = cost; bands = bnd; }
        }
    }
    a.band_rows = (LH + bands - 1) / bands;
    a.n_bands = (LH + a.band_rows - 1) / a.band_rows;
    const long blocks = (long)s.B * s.G * a.n_parts * a.n_bands;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    ++g_glr_stream_launches;
    GLR_PROF_BEGIN(slot, stream);
    int rc = -1000;
#ifndef GLRGTV_EMU
    // the geometries of the shipped v13 configuration (F = 6 / 12; planes of 256 .. 16 columns) have compile-time kernels
#define B2_TRY(LG_, NCH_) if (rc == -1000 && lg == LG_ && a.nch == NCH_) rc = b2_go<MODE, COARSE, LG_, NCH_, false>(a, blocks, threads, smem, stream);
    B2_TRY(8, 3) B2_TRY(7, 6) B2_TRY(6, 12) B2_TRY(6, 6) B2_TRY(5, 12) B2_TRY(4, 12)
#undef B2_TRY
#endif
    if (rc == -1000)
        rc = L > 32 ? b2_go<MODE, COARSE, 0, 0, true>(a, blocks, threads, smem, stream) : b2_go<MODE, COARSE, 0, 0, false>(a, blocks, threads, smem, stream);
    GLR_PROF_END(slot, stream);
    return rc ? rc : GLR_CHECK_LAUNCH();
}

// one backward stage: the half-resolution launch (writes vc), then the full-resolution launch
template <int MODE>
int glr_bw2_stage(B2Args a, const float* wT1, const float* wL1, float* gwT1, float* gwL1, float* vc, int slot, void* stream) {
    B2Args c = a;
    c.wT = wT1; c.wL = wL1; c.gwT = gwT1; c.gwL = gwL1; c.vc_out = vc; c.vc_in = nullptr;
    int rc = b2_launch<MODE, true>(c, slot, stream);
    if (rc) return rc;
    a.vc_in = vc; a.vc_out = nullptr;
    return b2_launch<MODE, false>(a, slot, stream);
}
