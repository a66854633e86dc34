// block_bwd.cu - fused backward of LocalLowpassFilteringBlock / MixtureGTVGLR.
//
// The reverse sweep mirrors the four forward stages.  Each kernel re-reads the saved stage input z
// (x2, x1, bA, y), RECOMPUTES the forward operator chain in shared memory (edge signals are never stored),
// runs the adjoint chain on the incoming gradient, and emits
//   - the gradient wrt z                                  (one tensor write),
//   - this stage's contribution to the four edge-weight gradients (accumulated over the graph's F
//     channels in registers, then one read-modify-write per tile),
//   - per-channel stats_kernel_p* gradients and per-graph scalar gradients (warp shuffles + shared atomics,
//     then one global atomic per value and CTA).
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
#include "stream_bwd.cuh"

enum { BWD_X3 = 0, BWD_X2 = 1, BWD_X1 = 2, BWD_BA = 3 };

// optional per-phase cycle accounting (debug builds only: tools/build_variant.sh x.so -DGLR_PHASE_TIMING)
#if defined(GLR_PHASE_TIMING) && !defined(GLRGTV_EMU)
__device__ unsigned long long g_bwd_phase[64];
#define PHASE_INIT() long long t_prev_ = clock64()
#define PHASE_MARK(slot)                                                                  \
    do {                                                                                  \
        if (threadIdx.x == 0) {                                                           \
            long long t_ = clock64();                                                     \
            atomicAdd(&g_bwd_phase[MODE * 8 + (slot)], (unsigned long long)(t_ - t_prev_)); \
            t_prev_ = t_;                                                                 \
        }                                                                                 \
    } while (0)
extern "C" int glrgtv_debug_bwd_phases(unsigned long long* out, int reset) {
    if (reset) { unsigned long long z[64] = {0}; return cudaMemcpyToSymbol(g_bwd_phase, z, sizeof(z)) == cudaSuccess ? 0 : -3; }
    return cudaMemcpyFromSymbol(out, g_bwd_phase, 64 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -3;
}
#else
#define PHASE_INIT() ((void)0)
#define PHASE_MARK(slot) ((void)0)
#endif

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

// shared-memory layout (floats) of one backward stage (4 guard floats on either side, see FwdLayout)
template <int MODE, int TH, int TW, int NT, bool GEN>
struct BwdLayout {
    using GF = Geo<TH, TW, NT, GEN>;
    using GC = Geo<TH / 2, TW / 2, NT, GEN>;
    static constexpr bool HAS_A = MODE != BWD_BA, HAS_R = MODE == BWD_X2 || MODE == BWD_BA, THR = MODE == BWD_X2;
    static constexpr int F3 = GF::floats(3), F2 = GF::floats(2), F1 = GF::floats(1);
    static constexpr int C3 = GC::floats(3), C2 = GC::floats(2), C1 = GC::floats(1);
    static constexpr int zf = 4;
    static constexpr int gA = zf + F3;
    static constexpr int gB = gA + (HAS_A ? F3 : 0);
    static constexpr int sA = gB + (HAS_R ? F3 : 0);
    static constexpr int sB = sA + (HAS_A ? F2 : 0);
    static constexpr int gl = sB + F2;
    static constexpr int goA = gl + (HAS_A ? F2 : 0);
    static constexpr int goB = goA + (HAS_A ? F2 : 0);
    static constexpr int lA = goB + (HAS_R ? F2 : 0);
    static constexpr int oB = lA + (HAS_A ? F1 : 0);
    static constexpr int oT = oB + F1;
    static constexpr int gsL = oT + (THR ? F1 : 0);
    static constexpr int gsT = gsL + (HAS_A ? F1 : 0);
    static constexpr int pz = gsT + F1;
    static constexpr int gcA = pz + C3;
    static constexpr int gcB = gcA + (HAS_A ? C3 : 0);
    static constexpr int sA1 = gcB + (HAS_R ? C3 : 0);
    static constexpr int sB1 = sA1 + (HAS_A ? C2 : 0);
    static constexpr int gl1 = sB1 + C2;
    static constexpr int goA1 = gl1 + (HAS_A ? C2 : 0);
    static constexpr int goB1 = goA1 + (HAS_A ? C2 : 0);
    static constexpr int lA1 = goB1 + (HAS_R ? C2 : 0);
    static constexpr int oB1 = lA1 + (HAS_A ? C1 : 0);
    static constexpr int oT1 = oB1 + C1;
    static constexpr int gsL1 = oT1 + (THR ? C1 : 0);
    static constexpr int gsT1 = gsL1 + (HAS_A ? C1 : 0);
    // weights: raw wL (the L adjoint reads the neighbours' weights -> halo 2); GTV: coefficient planes, or the
    // raw planes where the thresholded core needs them (then the linear core uses the raw planes too)
    static constexpr int wL0 = gsT1 + C1;
    static constexpr int cR0 = wL0 + (HAS_A ? 4 * F2 : 0);
    static constexpr int cD0 = cR0 + (THR ? 0 : F2);
    static constexpr int wT0 = cD0 + (THR ? 0 : F2);
    static constexpr int wL1 = wT0 + (THR ? 4 * F2 : 0);
    static constexpr int cR1 = wL1 + (HAS_A ? 4 * C2 : 0);
    static constexpr int cD1 = cR1 + (THR ? 0 : C2);
    static constexpr int wT1 = cD1 + (THR ? 0 : C2);
    static constexpr int red = wT1 + (THR ? 4 * C2 : 0) + 4;
    // cp.async staging of the next channel: z, and the one or two tensors the upstream gradients are built from
    static constexpr int raw_z = red + 64;
    static constexpr int raw_g0 = raw_z + Raw<GF>::FLOATS;
    static constexpr int raw_g1 = raw_g0 + Raw<GF>::FLOATS;
    static constexpr int total = raw_g1 + (MODE == BWD_X2 ? Raw<GF>::FLOATS : 0);
};

// per-thread accumulators that live across the channel loop: on the GPU every thread owns at most ONE
// epilogue quad, so a plain register array is enough; the emulation build (one thread per block) keeps one
// slot per quad in static storage.
#ifdef GLRGTV_EMU
#define ACC_DECL(name, items, n) static thread_local float name##_store[(items) * (n)]
#define ACC_PTR(name, i, n) (name##_store + (i) * (n))
#define EPI_LOOP(i, n, first) for (int i = 0; i < (n); ++i)
#else
// a thread is either a fine-quad owner or a coarse-quad owner, never both: one register array serves both roles
#define ACC_DECL(name, items, n)
#define ACC_PTR(name, i, n) (acc_regs)
#define EPI_LOOP(i, n, first) for (int i = (int)threadIdx.x - (first); i >= 0 && i < (n); i += (1 << 20))
#endif

// per-tap sums (c, R, D, U, L) of one stats kernel: acc[t] += g * v_t; they are turned into the four parameter
// gradients (p01 = c, p02a = R - c, p02b = D - c, p03 = 4c - R - D - U - L) once per channel at commit time
__device__ __forceinline__ void stats_acc(float* acc, float g, float vc, float vr, float vd, float vu, float vl) {
    acc[0] += g * vc;
    acc[1] += g * vr;
    acc[2] += g * vd;
    acc[3] += g * vu;
    acc[4] += g * vl;
}

// Parameter-gradient work of one epilogue quad at one resolution (fine: z halo 6, coarse: z halo 3).
//   st   : [T: c,R,D,U,L | L: c,R,D,U,L]    per-channel tap sums of this resolution's two modules
//   sums : [mu, ro, gamma]                  per-graph sums of this resolution
//   acc  : [L e0..e3][4] then [T e0..e3][4] edge-weight gradient accumulators of this quad
// returns the forward St values (glr, gtv_lin) and the S-adjoint values VT+VL in V.
template <int MODE, class G, int RZ>
struct QuadWork {
    static constexpr bool HAS_A = MODE != BWD_BA, HAS_R = MODE == BWD_X2 || MODE == BWD_BA, THR = MODE == BWD_X2;
    static constexpr int P = G::P;
    const G& g;
    Plane<G, RZ> z;
    Plane<G, 2> sA, sB, gl, goA, goB;
    Plane<G, 1> lA, oB, oT, gsL, gsT;
    StatsTaps kT, kL;
    float aT, aL, Gam;

    // r = row within the tile, c = local column of the quad; ga/gb = upstream quads (loaded by the caller)
    // raw GTV weights of one tile quad, [edge][pixel]; they do not depend on the channel
    __device__ static __forceinline__ void load_we(const G& g, const float* __restrict__ wT_global, int h, int w, float (&we)[4][4]) {
        const size_t HW = (size_t)g.H * g.W, o = (size_t)h * g.W + w;
        const bool full = (g.W & 3) == 0 && g.quad_inside(h, w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (full) ld4(wT_global + e * HW + o, we[e]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) we[e][j] = g.inside(h, w + j) ? wT_global[e * HW + o + j] : 0.f;
            }
        }
    }

    __device__ __forceinline__ void run(int r, int c, const float (&ga)[4], const float (&gb)[4], const float (&we)[4][4],
                                        float* st, float* sums, float* acc, float (&V)[4], float (&glr)[4],
                                        float (&gtv_lin)[4]) const {
        const int h = g.h0 + r, w = g.gw(c);
        N5 n;
        float gtv_R[4];
        // ---- GTV: forward value, St gradient, S adjoint + S gradient
        ld_n5<P>(oB.lrc(r + 1, c), n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gtv_lin[j] = kT.kc * n.c[j] + kT.kr * n.L(j) + kT.kd * n.u[j] + kT.ku * n.d[j] + kT.kl * n.Rr(j);
            const float gu = aT * ((HAS_A ? ga[j] : 0.f) + (HAS_R && !THR ? gb[j] : 0.f));
            stats_acc(st, gu, n.c[j], n.L(j), n.u[j], n.d[j], n.Rr(j));
            gtv_R[j] = gtv_lin[j];
        }
        if (THR) {
            ld_n5<P>(oT.lrc(r + 1, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gtv_R[j] = kT.kc * n.c[j] + kT.kr * n.L(j) + kT.kd * n.u[j] + kT.ku * n.d[j] + kT.kl * n.Rr(j);
                const float gu = aT * gb[j];
                stats_acc(st, gu, n.c[j], n.L(j), n.u[j], n.d[j], n.Rr(j));
            }
        }
        N5 nz;
        ld_n5<P>(z.lrc(r + RZ, c), nz);
        ld_n5<P>(gsT.lrc(r + 1, c), n);
        const bool edge = h == 0 || h == g.H - 1 || w == 0 || w + 4 >= g.W;   // only border quads collect replicated taps
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float self = 0.f;
            if (edge) {
                if (w + j == g.W - 1) self += kT.kr;
                if (h == g.H - 1) self += kT.kd;
                if (h == 0) self += kT.ku;
                if (w + j == 0) self += kT.kl;
            }
            V[j] = (kT.kc + self) * n.c[j] + kT.kr * n.L(j) + kT.kd * n.u[j] + kT.ku * n.d[j] + kT.kl * n.Rr(j);
            const float gs = n.c[j];
            stats_acc(st, gs, nz.c[j], nz.Rr(j), nz.d[j], nz.u[j], nz.L(j));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (HAS_A) sums[1] += aT * ga[j] * gtv_lin[j];
            if (HAS_R) sums[1] += aT * gb[j] * gtv_R[j];
        }
        // ---- GLR
        if (HAS_A) {
            ld_n5<P>(lA.lrc(r + 1, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                glr[j] = kL.kc * n.c[j] + kL.kr * n.L(j) + kL.kd * n.u[j] + kL.ku * n.d[j] + kL.kl * n.Rr(j);
                const float gu = aL * ga[j];
                stats_acc(st + 5, gu, n.c[j], n.L(j), n.u[j], n.d[j], n.Rr(j));
                sums[0] += gu * glr[j];
            }
            ld_n5<P>(gsL.lrc(r + 1, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float self = 0.f;
                if (edge) {
                    if (w + j == g.W - 1) self += kL.kr;
                    if (h == g.H - 1) self += kL.kd;
                    if (h == 0) self += kL.ku;
                    if (w + j == 0) self += kL.kl;
                }
                V[j] += (kL.kc + self) * n.c[j] + kL.kr * n.L(j) + kL.kd * n.u[j] + kL.ku * n.d[j] + kL.kl * n.Rr(j);
                const float gs = n.c[j];
                stats_acc(st + 5, gs, nz.c[j], nz.Rr(j), nz.d[j], nz.u[j], nz.L(j));
            }
            // edge weights of L: gw_e -= gl * sA[n_e]
            float glq[4];
            ld4(gl.lrc(r + 2, c), glq);
            ld_n5<P>(sA.lrc(r + 2, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[0 * 4 + j] -= glq[j] * n.u[j];
                acc[1 * 4 + j] -= glq[j] * n.L(j);
                acc[2 * 4 + j] -= glq[j] * n.Rr(j);
                acc[3 * 4 + j] -= glq[j] * n.d[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) glr[j] = 0.f;
        }
        // ---- edge weights of GTV: gw_e += D phi(t) + D w phi'(t) d  (linear: 2 w D d), D = go[p]-go[n], d = s[p]-s[n]
        N5 ns;
        ld_n5<P>(sB.lrc(r + 2, c), ns);
        float* accT = acc + 16;
        if (HAS_A) {
            ld_n5<P>(goA.lrc(r + 2, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                accT[0 * 4 + j] += 2.f * we[0][j] * (n.c[j] - n.u[j]) * (ns.c[j] - ns.u[j]);
                accT[1 * 4 + j] += 2.f * we[1][j] * (n.c[j] - n.L(j)) * (ns.c[j] - ns.L(j));
                accT[2 * 4 + j] += 2.f * we[2][j] * (n.c[j] - n.Rr(j)) * (ns.c[j] - ns.Rr(j));
                accT[3 * 4 + j] += 2.f * we[3][j] * (n.c[j] - n.d[j]) * (ns.c[j] - ns.d[j]);
            }
        }
        if (HAS_R) {
            ld_n5<P>(goB.lrc(r + 2, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float D[4] = {n.c[j] - n.u[j], n.c[j] - n.L(j), n.c[j] - n.Rr(j), n.c[j] - n.d[j]};
                const float d[4] = {ns.c[j] - ns.u[j], ns.c[j] - ns.L(j), ns.c[j] - ns.Rr(j), ns.c[j] - ns.d[j]};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (THR) {
                        const float t = we[e][j] * d[e];
                        accT[e * 4 + j] += D[e] * glr_phi(t, Gam) + D[e] * we[e][j] * glr_dphi(t, Gam) * d[e];
                        if (fabsf(t) > Gam) sums[2] += D[e] * we[e][j] * (t > 0.f ? -2.f : 2.f);
                    } else {
                        accT[e * 4 + j] += 2.f * we[e][j] * D[e] * d[e];
                    }
                }
            }
        }
    }
};

// warp-reduce N values and add lane 0's totals into shared accumulators
template <int N>
__device__ __forceinline__ void warp_commit(float* vals, float* sh) {
#ifdef GLRGTV_EMU
    for (int k = 0; k < N; ++k) sh[k] += vals[k];
#else
#pragma unroll
    for (int k = 0; k < N; ++k) {
        float v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&sh[k], v);
    }
#endif
}

template <int MODE, int TH, int TW, int NT, bool GEN>
__global__ void __launch_bounds__(NT) k_block_bwd_stage(BlockBwdArgs a) {
    GLR_SMEM_DECL(smem);
    using LY = BwdLayout<MODE, TH, TW, NT, GEN>;
    using GF = typename LY::GF;
    using GC = typename LY::GC;
    constexpr bool HAS_A = LY::HAS_A, HAS_R = LY::HAS_R, THR = LY::THR;
    constexpr int NQF = GF::TR * (GF::NQ - 2), NQC = GC::TR * (GC::NQ - 2);   // epilogue quads: fine / coarse
    static_assert(NT >= NQF + NQC, "one epilogue quad per thread");
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    GF gf; gf.H = H; gf.W = W; gf.h0 = (tile / tiles_w) * TH; gf.w0 = (tile % tiles_w) * TW;
    GC gc; gc.H = H / 2; gc.W = W / 2; gc.h0 = gf.h0 / 2; gc.w0 = gf.w0 / 2;
    const size_t HW = (size_t)H * W, HWc = HW / 4;
    const bool vec = (W & 3) == 0, vecc = (gc.W & 3) == 0;

    auto zf = plane_at<GF, 3>(smem, LY::zf);
    auto gA = plane_at<GF, 3>(smem, LY::gA);
    auto gB = plane_at<GF, 3>(smem, LY::gB);
    auto sA = plane_at<GF, 2>(smem, LY::sA);
    auto sB = plane_at<GF, 2>(smem, LY::sB);
    auto gl = plane_at<GF, 2>(smem, LY::gl);
    auto goA = plane_at<GF, 2>(smem, LY::goA);
    auto goB = plane_at<GF, 2>(smem, LY::goB);
    auto lA = plane_at<GF, 1>(smem, LY::lA);
    auto oB = plane_at<GF, 1>(smem, LY::oB);
    auto oT = plane_at<GF, 1>(smem, LY::oT);
    auto gsL = plane_at<GF, 1>(smem, LY::gsL);
    auto gsT = plane_at<GF, 1>(smem, LY::gsT);
    auto pz = plane_at<GC, 3>(smem, LY::pz);
    auto gcA = plane_at<GC, 3>(smem, LY::gcA);
    auto gcB = plane_at<GC, 3>(smem, LY::gcB);
    auto sA1 = plane_at<GC, 2>(smem, LY::sA1);
    auto sB1 = plane_at<GC, 2>(smem, LY::sB1);
    auto gl1 = plane_at<GC, 2>(smem, LY::gl1);
    auto goA1 = plane_at<GC, 2>(smem, LY::goA1);
    auto goB1 = plane_at<GC, 2>(smem, LY::goB1);
    auto lA1 = plane_at<GC, 1>(smem, LY::lA1);
    auto oB1 = plane_at<GC, 1>(smem, LY::oB1);
    auto oT1 = plane_at<GC, 1>(smem, LY::oT1);
    auto gsL1 = plane_at<GC, 1>(smem, LY::gsL1);
    auto gsT1 = plane_at<GC, 1>(smem, LY::gsT1);
    auto wL0 = wplanes_at<GF, 2>(smem, LY::wL0);
    auto cR0 = plane_at<GF, 2>(smem, LY::cR0);
    auto cD0 = plane_at<GF, 2>(smem, LY::cD0);
    auto wT0 = wplanes_at<GF, 2>(smem, LY::wT0);
    auto wL1 = wplanes_at<GC, 2>(smem, LY::wL1);
    auto cR1 = plane_at<GC, 2>(smem, LY::cR1);
    auto cD1 = plane_at<GC, 2>(smem, LY::cD1);
    auto wT1 = wplanes_at<GC, 2>(smem, LY::wT1);
    float* red = smem + LY::red;  // [0:20) per-channel tap sums, [32:42) per-graph sums

    // ---- per-graph scalars
    const float aT0 = expf(a.p.ro0[g]), aT1 = expf(a.p.ro1[g]);
    const float aL0 = expf(a.p.mu0[g]), aL1 = expf(a.p.mu1[g]);
    const float G0 = expf(a.p.gamma0[g]), G1 = expf(a.p.gamma1[g]);
    const float al0 = a.p.alpha[g], al1 = a.p.alpha[G + g], al2 = a.p.alpha[2 * G + g], be2 = a.p.beta[2 * G + g];
    const bool has_skip = a.p.skip != nullptr;
    const float s0 = has_skip ? a.p.skip[0] : 0.f, s1 = has_skip ? a.p.skip[1] : 1.f;
    const float c23 = al2 * s1;  // gr2 = c23 * gout

    // branch-free borders: the zero-extended planes are zeroed once, outside quads are never written afterwards
    if (!GEN) {
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        TILE_LOOP_NT(NT, i, (LY::wL0 - LY::zf) / 4) st4(smem + LY::zf + 4 * i, z4);
    }

    // ---- weights
    const size_t wplane = (size_t)plane * 4;
    if (THR) {
        load_weights(gf, wT0, a.wT0 + wplane * HW);
        load_weights(gc, wT1, a.wT1 + wplane * HWc);
    } else {
        load_gtv_coeffs(gf, cR0, cD0, a.wT0 + wplane * HW);
        load_gtv_coeffs(gc, cR1, cD1, a.wT1 + wplane * HWc);
    }
    if (HAS_A) {
        load_weights(gf, wL0, a.wL0 + wplane * HW);
        load_weights(gc, wL1, a.wL1 + wplane * HWc);
    }
    TILE_LOOP_NT(NT, i, 64) red[i] = 0.f;

    // edge-weight gradient accumulators of this thread's epilogue quad: [L e][4] + [T e][4]
    ACC_DECL(accF, NQF, 32);
    ACC_DECL(accC, NQC, 32);
#ifdef GLRGTV_EMU
    for (int i = 0; i < NQF * 32; ++i) accF_store[i] = 0.f;
    for (int i = 0; i < NQC * 32; ++i) accC_store[i] = 0.f;
#else
    float acc_regs[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc_regs[i] = 0.f;
#endif
    // raw GTV weights of this thread's epilogue quad (fine owner or coarse owner): loaded once, reused by every channel
    float we[4][4];
#ifndef GLRGTV_EMU
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < 4; ++j) we[e][j] = 0.f;
    EPI_LOOP(i, NQF, 0) {
        const Quad qd = tile_quad_of(gf, i);
        if (qd.h < H && qd.w < W) QuadWork<MODE, GF, 3>::load_we(gf, a.wT0 + wplane * HW, qd.h, qd.w, we);
    }
    EPI_LOOP(i, NQC, NQF) {
        const Quad qd = tile_quad_of(gc, i);
        if (qd.h < gc.H && qd.w < gc.W) QuadWork<MODE, GC, 3>::load_we(gc, a.wT1 + wplane * HWc, qd.h, qd.w, we);
    }
#endif
    // per-graph sums of this thread: fine [mu0, ro0, gamma0, alpha_k, beta2, skip0, skip1], coarse [mu1, ro1, gamma1]
    float gsF[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gsC[3] = {0.f, 0.f, 0.f};

    float* rz = smem + LY::raw_z;
    float* rg0 = smem + LY::raw_g0;
    float* rg1 = smem + LY::raw_g1;
    // which global tensors feed the upstream gradients of this stage: X3: gout | X2: gout, gx2 | X1: gx1 | BA: gbA
    const float* src_g0 = (MODE == BWD_X3 || MODE == BWD_X2) ? a.gout : a.gin;
    const float* src_g1 = MODE == BWD_X2 ? a.gin : nullptr;
    const size_t plane0 = ((size_t)b * G * F + (size_t)g * F) * HW;
    auto stage_channel = [&](size_t o) {
        async_stage_raw(gf, rz, a.z + o);
        async_stage_raw(gf, rg0, src_g0 + o);
        if (MODE == BWD_X2) async_stage_raw(gf, rg1, src_g1 + o);
        cp_async_commit();
    };
    stage_channel(plane0);
    PHASE_INIT();
    PHASE_MARK(0);   // prologue: weights, staging of channel 0

    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const size_t off = plane0 + (size_t)f * HW;
        const StatsTaps kT0 = glr_load_taps(a.p.gtv0.stats, c), kT1 = glr_load_taps(a.p.gtv1.stats, c);
        const StatsTaps kL0 = glr_load_taps(a.p.glr0.stats, c), kL1 = glr_load_taps(a.p.glr1.stats, c);

        cp_async_wait_all();
        __syncthreads();
        PHASE_MARK(1);   // wait for the staged input / previous epilogue tail
        // ---- phase 0: stage input (clamp-extended) and upstream gradients (zero-extended), fine (+)3 and pooled
        consume_raw<true>(gf, gc, zf, pz, rz, (const float*)nullptr, [](float v, float) { return v; });
        if (MODE == BWD_X3) {
            consume_raw<false>(gf, gc, gA, gcA, rg0, (const float*)nullptr, [=](float go, float) { return -c23 * go; });
        } else if (MODE == BWD_X2) {
            consume_raw<false>(gf, gc, gA, gcA, rg0, rg1, [=](float go, float gx2) { return -(be2 * c23 * go + al1 * gx2); });
            consume_raw<false>(gf, gc, gB, gcB, rg0, rg1, [=](float go, float gx2) { return c23 * go + (be2 * c23 * go + al1 * gx2); });
        } else if (MODE == BWD_X1) {
            consume_raw<false>(gf, gc, gA, gcA, rg0, (const float*)nullptr, [=](float gx1, float) { return -al0 * gx1; });
        } else {
            consume_raw<false>(gf, gc, gB, gcB, rg0, (const float*)nullptr, [](float v, float) { return v; });
        }
        __syncthreads();
        PHASE_MARK(2);   // consume
        if (f + 1 < F) stage_channel(off + HW);
        // ---- phase 1: forward S and the St-adjoints of the upstreams, both resolutions
        TILE_LOOP_NT(NT, i, GF::items(2)) {
            const Quad q = quad_of<GF, 2>(gf, i);
            q_S<HAS_A>(gf, q, HAS_A ? sA : sB, HAS_A ? kL0 : kT0, sB, kT0, zf);
            if (HAS_A) q_Szero<true, true>(gf, q, gl, kL0, aL0, goA, kT0, aT0, gA);
            if (HAS_R) q_Szero<false, true>(gf, q, goB, kT0, 0.f, goB, kT0, aT0, gB);
        }
        TILE_LOOP_REV(NT, i, GC::items(2)) {
            const Quad q = quad_of<GC, 2>(gc, i);
            q_S<HAS_A>(gc, q, HAS_A ? sA1 : sB1, HAS_A ? kL1 : kT1, sB1, kT1, pz);
            if (HAS_A) q_Szero<true, true>(gc, q, gl1, kL1, aL1, goA1, kT1, aT1, gcA);
            if (HAS_R) q_Szero<false, true>(gc, q, goB1, kT1, 0.f, goB1, kT1, aT1, gcB);
        }
        __syncthreads();
        PHASE_MARK(3);   // phase 1
        // ---- phase 2: forward cores and adjoint cores, both resolutions
        TILE_LOOP_NT(NT, i, GF::items(1)) {
            const Quad q = quad_of<GF, 1>(gf, i);
            if (HAS_A) { q_L(gf, q, lA, sA, wL0); q_L_adj(gf, q, gsL, gl, wL0); }
            if (THR) {
                q_gtv_raw<true, true>(gf, q, oB, oT, sB, wT0, G0);
                q_gtv_raw_adj<true, true>(gf, q, gsT, goA, goB, sB, wT0, G0);
            } else {
                q_gtv_lin(gf, q, oB, sB, cR0, cD0);
                q_gtv_lin(gf, q, gsT, HAS_A ? goA : goB, cR0, cD0);   // the linear core is self-adjoint
            }
        }
        TILE_LOOP_REV(NT, i, GC::items(1)) {
            const Quad q = quad_of<GC, 1>(gc, i);
            if (HAS_A) { q_L(gc, q, lA1, sA1, wL1); q_L_adj(gc, q, gsL1, gl1, wL1); }
            if (THR) {
                q_gtv_raw<true, true>(gc, q, oB1, oT1, sB1, wT1, G1);
                q_gtv_raw_adj<true, true>(gc, q, gsT1, goA1, goB1, sB1, wT1, G1);
            } else {
                q_gtv_lin(gc, q, oB1, sB1, cR1, cD1);
                q_gtv_lin(gc, q, gsT1, HAS_A ? goA1 : goB1, cR1, cD1);
            }
        }
        __syncthreads();
        PHASE_MARK(4);   // phase 2
        // ---- phase 3: epilogues.  Threads [0, NQF) own one fine quad each, threads [NQF, NQF+NQC) one coarse quad.
        float stF[10], stC[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) stF[k] = stC[k] = 0.f;
        EPI_LOOP(i, NQC, NQF) {
            const Quad qd = tile_quad_of(gc, i);
            const int r = qd.r, cq = qd.c, h = qd.h, w = qd.w;
            if (h >= gc.H || w >= gc.W) continue;
            QuadWork<MODE, GC, 3> qw{gc, pz, sA1, sB1, gl1, goA1, goB1, lA1, oB1, oT1, gsL1, gsT1, kT1, kL1, aT1, aL1, G1};
#ifdef GLRGTV_EMU
            QuadWork<MODE, GC, 3>::load_we(gc, a.wT1 + wplane * HWc, h, w, we);
#endif
            float ga[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f}, V[4], glr[4], gtvl[4];
            if (HAS_A) ld4(gcA.lrc(r + 3, cq), ga);
            if (HAS_R) ld4(gcB.lrc(r + 3, cq), gb);
            qw.run(r, cq, ga, gb, we, stC, gsC, ACC_PTR(accC, i, 32), V, glr, gtvl);
        }
        EPI_LOOP(i, NQF, 0) {
            const Quad qd = tile_quad_of(gf, i);
            const int r = qd.r, cq = qd.c, h = qd.h, w = qd.w;
            if (h >= H || w >= W) continue;
            QuadWork<MODE, GF, 3> qw{gf, zf, sA, sB, gl, goA, goB, lA, oB, oT, gsL, gsT, kT0, kL0, aT0, aL0, G0};
#ifdef GLRGTV_EMU
            QuadWork<MODE, GF, 3>::load_we(gf, a.wT0 + wplane * HW, h, w, we);
#endif
            // pointwise operands from global memory first, so that their latency hides behind the shared-memory work:
            // q0 gout, q1 gin, q2 r1 / gx2, q3 bB, q4 x
            const size_t gi = off + (size_t)h * W + w;
            const bool full = !GEN || (vec && w + 3 < W);
            float q0[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f}, q2[4] = {0.f, 0.f, 0.f, 0.f},
                  q3[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f}, outv[4];
            if (full) {
                if (MODE == BWD_X3 || MODE == BWD_X2 || MODE == BWD_BA) ld4(a.gout + gi, q0);
                if (MODE != BWD_X3) ld4(a.gin + gi, q1);
                if (MODE == BWD_X3 || MODE == BWD_X2) ld4(a.r1 + gi, q2);
                if (MODE == BWD_BA) ld4(a.gx2 + gi, q2);
                if (MODE == BWD_X3) { ld4(a.bB + gi, q3); if (has_skip) ld4(a.x + gi, q4); }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (w + j >= W) continue;
                    if (MODE == BWD_X3 || MODE == BWD_X2 || MODE == BWD_BA) q0[j] = a.gout[gi + j];
                    if (MODE != BWD_X3) q1[j] = a.gin[gi + j];
                    if (MODE == BWD_X3 || MODE == BWD_X2) q2[j] = a.r1[gi + j];
                    if (MODE == BWD_BA) q2[j] = a.gx2[gi + j];
                    if (MODE == BWD_X3) { q3[j] = a.bB[gi + j]; if (has_skip) q4[j] = a.x[gi + j]; }
                }
            }
            float ga[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f}, V[4], glr[4], gtvl[4];
            if (HAS_A) ld4(gA.lrc(r + 3, cq), ga);
            if (HAS_R) ld4(gB.lrc(r + 3, cq), gb);
            qw.run(r, cq, ga, gb, we, stF, gsF, ACC_PTR(accF, i, 32), V, glr, gtvl);
            // gradient coming back through the coarse branch (VJP of P is P^T: 0.25 * replicate), inline per coarse pixel
            const int rc = (r >> 1) + 1, cc = (cq >> 1) + 2, hc = h >> 1, wc = (w >> 1);
            float gz0 = S_adj_elem(gsT1.lrc(rc, cc), GC::P, kT1, hc, wc, gc.H, gc.W);
            float gz1 = S_adj_elem(gsT1.lrc(rc, cc + 1), GC::P, kT1, hc, wc + 1, gc.H, gc.W);
            if (HAS_A) {
                gz0 += S_adj_elem(gsL1.lrc(rc, cc), GC::P, kL1, hc, wc, gc.H, gc.W);
                gz1 += S_adj_elem(gsL1.lrc(rc, cc + 1), GC::P, kL1, hc, wc + 1, gc.H, gc.W);
            }
            float zq[4];
            ld4(zf.lrc(r + 3, cq), zq);
            // forward A(z) where the stage needs it (coarse forward term inline, as in the forward kernel)
            float Az[4] = {0.f, 0.f, 0.f, 0.f};
            if (MODE == BWD_X3 || MODE == BWD_X1) {
                const float t0 = aT1 * St_elem(oB1.lrc(rc, cc), GC::P, kT1) + aL1 * St_elem(lA1.lrc(rc, cc), GC::P, kL1);
                const float t1 = aT1 * St_elem(oB1.lrc(rc, cc + 1), GC::P, kT1) + aL1 * St_elem(lA1.lrc(rc, cc + 1), GC::P, kL1);
#pragma unroll
                for (int j = 0; j < 4; ++j) Az[j] = zq[j] + aL0 * glr[j] + aT0 * gtvl[j] + 0.25f * (j < 2 ? t0 : t1);
            }
            // (pointwise operands q0 gout, q1 gin, q2 r1 / gx2, q3 bB, q4 x were loaded at the top of the item)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float Vj = (HAS_A ? ga[j] : 0.f) + V[j] + 0.25f * (j < 2 ? gz0 : gz1);
                if (MODE == BWD_X3) {
                    const float u2 = (q3[j] - Az[j]) + be2 * q2[j], x3 = zq[j] + al2 * u2, g3 = s1 * q0[j];
                    outv[j] = g3 + Vj;
                    gsF[3] += g3 * u2;
                    gsF[4] += al2 * g3 * q2[j];
                    if (has_skip) { gsF[5] += q0[j] * q4[j]; gsF[6] += q0[j] * x3; }
                } else if (MODE == BWD_X2) {
                    outv[j] = q1[j] + Vj;
                    gsF[3] += q1[j] * q2[j];
                } else if (MODE == BWD_X1) {
                    outv[j] = (1.f + al0) * q1[j] + Vj;
                    gsF[3] += q1[j] * (zq[j] - Az[j]);
                } else {
                    const float gr2 = c23 * q0[j];
                    outv[j] = q1[j] + Vj + (gr2 + be2 * gr2 + al1 * q2[j]) + s0 * q0[j];
                }
            }
            if (full) st4(a.gz_out + gi, outv);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w + j < W) a.gz_out[gi + j] = outv[j];
            }
        }
        // ---- per-channel stats gradients: warp shuffles -> shared atomics -> one global atomic per value
        warp_commit<10>(stF, red);
        warp_commit<10>(stC, red + 10);
        __syncthreads();
        PHASE_MARK(5);   // epilogues + stats commit
        {
            float* dst[4] = {a.gr.gtv0_stats, a.gr.glr0_stats, a.gr.gtv1_stats, a.gr.glr1_stats};
            const int C = G * F;
            TILE_LOOP_NT(NT, t, 16) {   // red: [T0 | L0 | T1 | L1] x (c, R, D, U, L)
                const int m = t >> 2, k = t & 3;
                const float* q = red + 5 * m;
                const float v = k == 0 ? q[0] : k == 1 ? q[1] - q[0] : k == 2 ? q[2] - q[0] : 4.f * q[0] - q[1] - q[2] - q[3] - q[4];
                if (HAS_A || m == 0 || m == 2) atomicAdd(&dst[m][k * C + c], v);
            }
        }
        __syncthreads();
        TILE_LOOP_NT(NT, t, 20) red[t] = 0.f;
        PHASE_MARK(6);   // stats atomics
    }

    // ---- edge-weight gradients of this tile: one read-modify-write per stage
    {
        float* gwF[2] = {a.gwL0 + wplane * HW, a.gwT0 + wplane * HW};
        EPI_LOOP(i, NQF, 0) {
            const Quad qd = tile_quad_of(gf, i);
            const int h = qd.h, w = qd.w;
            if (h >= H || w >= W) continue;
            const float* acc = ACC_PTR(accF, i, 32);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (!HAS_A && m == 0) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float* q = gwF[m] + (size_t)e * HW + (size_t)h * W + w;
                    float v[4] = {acc[m * 16 + e * 4 + 0], acc[m * 16 + e * 4 + 1], acc[m * 16 + e * 4 + 2], acc[m * 16 + e * 4 + 3]};
                    if (!GEN || (vec && w + 3 < W)) {
                        if (!a.gw_assign) { float o[4]; ld4(q, o); v[0] += o[0]; v[1] += o[1]; v[2] += o[2]; v[3] += o[3]; }
                        st4(q, v);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (w + j < W) q[j] = a.gw_assign ? v[j] : q[j] + v[j];
                    }
                }
            }
        }
        float* gwC[2] = {a.gwL1 + wplane * HWc, a.gwT1 + wplane * HWc};
        EPI_LOOP(i, NQC, NQF) {
            const Quad qd = tile_quad_of(gc, i);
            const int h = qd.h, w = qd.w;
            if (h >= gc.H || w >= gc.W) continue;
            const float* acc = ACC_PTR(accC, i, 32);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (!HAS_A && m == 0) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float* q = gwC[m] + (size_t)e * HWc + (size_t)h * gc.W + w;
                    float v[4] = {acc[m * 16 + e * 4 + 0], acc[m * 16 + e * 4 + 1], acc[m * 16 + e * 4 + 2], acc[m * 16 + e * 4 + 3]};
                    if (!GEN || (vecc && w + 3 < gc.W)) {
                        if (!a.gw_assign) { float o[4]; ld4(q, o); v[0] += o[0]; v[1] += o[1]; v[2] += o[2]; v[3] += o[3]; }
                        st4(q, v);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (w + j < gc.W) q[j] = a.gw_assign ? v[j] : q[j] + v[j];
                    }
                }
            }
        }
    }
    // ---- per-graph scalar gradients
    __syncthreads();
    warp_commit<7>(gsF, red + 32);
    warp_commit<3>(gsC, red + 39);
    __syncthreads();
    if (threadIdx.x == 0) {
        const float* s = red + 32;  // mu0, ro0, gamma0, alpha_k, beta2, skip0, skip1, mu1, ro1, gamma1
        if (HAS_A) {
            atomicAdd(&a.gr.mu0[g], s[0]);
            atomicAdd(&a.gr.mu1[g], s[7]);
        }
        atomicAdd(&a.gr.ro0[g], s[1]);
        atomicAdd(&a.gr.ro1[g], s[8]);
        if (THR) {
            atomicAdd(&a.gr.gamma0[g], s[2] * G0);
            atomicAdd(&a.gr.gamma1[g], s[9] * G1);
        }
        if (MODE == BWD_X3) {
            atomicAdd(&a.gr.alpha[2 * G + g], s[3]);
            atomicAdd(&a.gr.beta[2 * G + g], s[4]);
            if (has_skip && a.gr.skip) {
                atomicAdd(&a.gr.skip[0], s[5]);
                atomicAdd(&a.gr.skip[1], s[6]);
            }
        }
        if (MODE == BWD_X2) atomicAdd(&a.gr.alpha[G + g], s[3]);
        if (MODE == BWD_X1) atomicAdd(&a.gr.alpha[g], s[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#ifndef GLR_BTH
#define GLR_BTH 32
#define GLR_BTW 32
#endif
#ifndef GLR_BWD_THREADS
#define GLR_BWD_THREADS 384
#endif

template <int MODE, bool GEN>
static int launch_bwd_stage_gen(const BlockBwdArgs& a, void* stream) {
    const glrgtv_shape& s = a.s;
    const long tiles = (long)((s.W + GLR_BTW - 1) / GLR_BTW) * ((s.H + GLR_BTH - 1) / GLR_BTH);
    const long blocks = tiles * s.B * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    constexpr size_t smem = (size_t)BwdLayout<MODE, GLR_BTH, GLR_BTW, GLR_BWD_THREADS, GEN>::total * sizeof(float);
    static_assert(smem <= 227 * 1024, "backward tile does not fit shared memory");
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_block_bwd_stage<MODE, GLR_BTH, GLR_BTW, GLR_BWD_THREADS, GEN>, smem, optin)) return rc_;
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_X3 + MODE, stream);
    GLR_LAUNCH((k_block_bwd_stage<MODE, GLR_BTH, GLR_BTW, GLR_BWD_THREADS, GEN>), dim3((unsigned)blocks), GLR_BWD_THREADS,
               smem, stream, a);
    GLR_PROF_END(GLRGTV_SLOT_BWD_X3 + MODE, stream);
    return GLR_CHECK_LAUNCH();
}
template <int MODE>
static int launch_bwd_stage(const BlockBwdArgs& a, void* stream) {
    return (a.s.W % 8 == 0) ? launch_bwd_stage_gen<MODE, false>(a, stream) : launch_bwd_stage_gen<MODE, true>(a, stream);
}

// edge-weight gradients of one stage of the round-1 backward: the tiled kernel of block_gw.cu
template <int MODE>
static int gw_stage(const GwArgs& w, int slot, void* stream) { return glr_gw_stage<MODE>(w, slot, stream); }

// tiled edge-weight backward of one resolution (block_weights_bwd.cu)
int glr_block_weights_bwd(const glrgtv_shape* s, const float* feat, const float* M_gtv, const float* M_glr,
                          const float* w_gtv, const float* w_glr, const float* gw_gtv, const float* gw_glr, float* gfeat,
                          float* gM_gtv, float* gM_glr, void* stream);

// workspace layout in floats; every segment starts 16-byte aligned
static size_t ws_floats(const glrgtv_shape* s, size_t* o_gx2, size_t* o_gx1, size_t* o_gbA, size_t* o_gw, size_t* o_scr) {
    const size_t N = (size_t)s->B * s->H * s->W, C = (size_t)s->G * s->F, GE = (size_t)s->G * 4;
    auto up = [](size_t v) { return (v + 3) & ~(size_t)3; };
    size_t off = 0;
    *o_gx2 = off; off = up(off + C * N);
    *o_gx1 = off; off = up(off + C * N);
    *o_gbA = off; off = up(off + C * N);
    *o_gw = off;  off = up(off + 2 * GE * N + 2 * GE * (N / 4));
    *o_scr = off;
    return off;
}

// second-generation backward (bw2.cu): five stage tensors (gx2, gA, gB, gx1, gbA), the half-resolution result vc, and the four
// edge-weight gradient sets (accumulated by the stages with red.global.add: zeroed first)
struct Bw2Ws {
    size_t gx2, gA, gB, gx1, gbA, vc, gw, total;
};
static Bw2Ws bw2_ws(const glrgtv_shape* s) {
    const size_t N = (size_t)s->B * s->H * s->W, C = (size_t)s->G * s->F, GE = (size_t)s->G * 4;
    auto up = [](size_t v) { return (v + 3) & ~(size_t)3; };
    Bw2Ws w;
    size_t off = 0;
    w.gx2 = off; off = up(off + C * N);
    w.gA = off; off = up(off + C * N);
    w.gB = off; off = up(off + C * N);
    w.gx1 = off; off = up(off + C * N);
    w.gbA = off; off = up(off + C * N);
    w.vc = off; off = up(off + C * (N / 4));
    w.gw = off; off = up(off + 2 * GE * N + 2 * GE * (N / 4));
    w.total = off;
    return w;
}
bool glr_bw2_eligible(const glrgtv_shape* s);
extern int g_glr_bw2;

extern "C" size_t glrgtv_block_bwd_workspace_bytes(const glrgtv_shape* s) {
    if (!glr_shape_ok(s)) return 0;
    size_t a, b, c, d, e;
    size_t n = ws_floats(s, &a, &b, &c, &d, &e);
    if (glr_bw2_eligible(s)) { const size_t m = bw2_ws(s).total; n = m > n ? m : n; }
    return n * sizeof(float);
}

int glr_block_params_ok(const glrgtv_shape* s, const glrgtv_block_params* p);
// block_stream_fwd.cu
extern int g_glr_block_path;
int glr_stream_fwd_eligible(const glrgtv_shape* s);
#ifndef GLR_STREAM_BWD_MIN_W
#define GLR_STREAM_BWD_MIN_W 8      // narrower planes take the plane kernels in automatic mode
#endif

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
    for (int i = 0; i < 9; ++i) {
        GLR_REQUIRE_PTR(need[i]);
        if (!glr_aligned16(need[i])) return GLRGTV_ERR_POINTER;
    }
    if (!glr_aligned16(x) || !glr_aligned16(gout) || !glr_aligned16(gx) || !glr_aligned16(workspace)) return GLRGTV_ERR_POINTER;
    float* const* gp = &gr->gtv0_stats;
    for (int i = 0; i < 16; ++i) GLR_REQUIRE_PTR(gp[i]);
    if (p->skip) GLR_REQUIRE_PTR(gr->skip);
    size_t o_gx2, o_gx1, o_gbA, o_gw, o_scr;
    if (workspace_bytes < ws_floats(s, &o_gx2, &o_gx1, &o_gbA, &o_gw, &o_scr) * sizeof(float)) return GLRGTV_ERR_WORKSPACE;

    float* ws = (float*)workspace;
    const size_t N = (size_t)s->B * s->H * s->W, GE = (size_t)s->G * 4;
    // second-generation pair walkers (bw2.cu): edge-weight gradients folded in, 2 launches per stage (half, full resolution)
    if (g_glr_bw2 && g_glr_block_path != 1 && glr_bw2_eligible(s)) {
        const Bw2Ws o = bw2_ws(s);
        if (workspace_bytes < o.total * sizeof(float)) return GLRGTV_ERR_WORKSPACE;
        float *gwT0 = ws + o.gw, *gwL0 = gwT0 + GE * N, *gwT1 = gwL0 + GE * N, *gwL1 = gwT1 + GE * (N / 4);
        if (glr_memset_async(gwT0, 0, (2 * GE * N + 2 * GE * (N / 4)) * sizeof(float), (cudaStream_t)stream)) return GLRGTV_ERR_CUDA;
        B2Args b = {};
        b.s = *s; b.p = *p; b.gr = *gr;
        b.wT = sv->wT0; b.wL = sv->wL0; b.gwT = gwT0; b.gwL = gwL0;
        float *gx2 = ws + o.gx2, *gA = ws + o.gA, *gB = ws + o.gB, *gx1 = ws + o.gx1, *gbA = ws + o.gbA, *vc = ws + o.vc;
        // X3: through x3 = x2 + a2 (r2 + b2 r1), r2 = bB - A x2
        b.z = sv->x2; b.src = gout; b.op0 = sv->r1; b.op1 = sv->bB; b.op2 = x; b.out0 = gx2; b.out1 = gA; b.out2 = gB;
        if ((rc = glr_bw2_stage<BW_X3>(b, sv->wT1, sv->wL1, gwT1, gwL1, vc, GLRGTV_SLOT_BWD_X3, stream))) return rc;
        // X2: through x2 = x1 + a1 r1, r1 = bB - A x1 (part A) and bB = y + R_thr x1 (part B)
        b.z = sv->x1; b.src = gA; b.op0 = sv->r1; b.op1 = gx2; b.op2 = nullptr; b.out0 = gx1; b.out1 = b.out2 = nullptr;
        if ((rc = glr_bw2_stage<BW_X2A>(b, sv->wT1, sv->wL1, gwT1, gwL1, vc, GLRGTV_SLOT_BWD_X2, stream))) return rc;
        b.src = gB; b.op0 = gx1; b.op1 = nullptr;
        if ((rc = glr_bw2_stage<BW_X2B>(b, sv->wT1, sv->wL1, gwT1, gwL1, vc, GLRGTV_SLOT_BWD_X2, stream))) return rc;
        // X1: through x1 = bA + a0 (bA - A bA)
        b.z = sv->bA; b.src = gx1; b.op0 = nullptr; b.out0 = gbA;
        if ((rc = glr_bw2_stage<BW_X1>(b, sv->wT1, sv->wL1, gwT1, gwL1, vc, GLRGTV_SLOT_BWD_X1, stream))) return rc;
        // BA: through bA = y + R_lin y, plus the pointwise paths into y (bB, skip)
        b.z = x; b.src = gbA; b.op0 = gB; b.op1 = gout; b.out0 = gx;
        if ((rc = glr_bw2_stage<BW_BA>(b, sv->wT1, sv->wL1, gwT1, gwL1, vc, GLRGTV_SLOT_BWD_BA, stream))) return rc;
        glrgtv_shape sc = *s;
        sc.H /= 2; sc.W /= 2;
        GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_WEIGHTS, stream);
        if ((rc = glr_block_weights_bwd(s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, gwT0, gwL0, gfeat0, gr->gtv0_M,
                                        gr->glr0_M, stream))) return rc;
        rc = glr_block_weights_bwd(&sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, gwT1, gwL1, gfeat1, gr->gtv1_M,
                                   gr->glr1_M, stream);
        GLR_PROF_END(GLRGTV_SLOT_BWD_WEIGHTS, stream);
        return rc;
    }
    // register-streaming backward (block_stream_bwd.cu + block_gw.cu) where the shape allows, else the plane kernels
    const bool can_stream = glr_stream_fwd_eligible(s) && sv->cT0 && sv->cT1;     // any W % 8 == 0: wide planes in column strips
    if (g_glr_block_path == 2 && !can_stream) return GLRGTV_ERR_UNSUPPORTED;
    if (can_stream && g_glr_block_path != 1 && (g_glr_block_path == 2 || s->W >= GLR_STREAM_BWD_MIN_W)) {
        StreamBwdArgs b;
        b.s = *s; b.p = *p; b.gr = *gr;
        b.wT0 = sv->wT0; b.wL0 = sv->wL0; b.wT1 = sv->wT1; b.wL1 = sv->wL1; b.cT0 = sv->cT0; b.cT1 = sv->cT1;
        b.nch = 1; b.band_rows = s->H; b.n_bands = 1; b.n_strips = 1;
        GwArgs w;
        w.s = *s; w.p = *p; w.ggamma0 = gr->gamma0; w.ggamma1 = gr->gamma1; w.wT0 = sv->wT0; w.wT1 = sv->wT1;
        w.gwT0 = ws + o_gw; w.gwL0 = w.gwT0 + GE * N; w.gwT1 = w.gwL0 + GE * N; w.gwL1 = w.gwT1 + GE * (N / 4);
        float *gx2 = ws + o_gx2, *gx1 = ws + o_gx1, *gbA = ws + o_gbA;
        // X3: through x3 = x2 + a2 (r2 + b2 r1), r2 = bB - A x2
        b.z = sv->x2; b.src0 = gout; b.src1 = nullptr; b.op0 = sv->r1; b.op1 = sv->bB; b.op2 = x; b.out = gx2;
        if ((rc = glr_stream_bwd_stage<BW_X3>(b, GLRGTV_SLOT_BWD_X3, stream))) return rc;
        w.z = sv->x2; w.src0 = gout; w.src1 = nullptr; w.assign = 1;
        if ((rc = gw_stage<BW_X3>(w, GLRGTV_SLOT_GW, stream))) return rc;
        // X2: through x2 = x1 + a1 r1, r1 = bB - A x1 (part A) and bB = y + R_thr x1 (part B)
        b.z = sv->x1; b.src0 = gout; b.src1 = gx2; b.op0 = sv->r1; b.op1 = nullptr; b.op2 = nullptr; b.out = gx1;
        if ((rc = glr_stream_bwd_stage<BW_X2A>(b, GLRGTV_SLOT_BWD_X2, stream))) return rc;
        b.op0 = gx1;
        if ((rc = glr_stream_bwd_stage<BW_X2B>(b, GLRGTV_SLOT_BWD_X2, stream))) return rc;
        w.z = sv->x1; w.src0 = gout; w.src1 = gx2; w.assign = 0;
        if ((rc = gw_stage<BW_X2A>(w, GLRGTV_SLOT_GW, stream))) return rc;      // both parts of X2 in one pass
        // X1: through x1 = bA + a0 (bA - A bA)
        b.z = sv->bA; b.src0 = gx1; b.src1 = nullptr; b.op0 = nullptr; b.out = gbA;
        if ((rc = glr_stream_bwd_stage<BW_X1>(b, GLRGTV_SLOT_BWD_X1, stream))) return rc;
        w.z = sv->bA; w.src0 = gx1; w.src1 = nullptr;
        if ((rc = gw_stage<BW_X1>(w, GLRGTV_SLOT_GW, stream))) return rc;
        // BA: through bA = y + R_lin y, plus the pointwise paths into y (bB, skip)
        b.z = x; b.src0 = gbA; b.op0 = gout; b.op1 = gx2; b.out = gx;
        if ((rc = glr_stream_bwd_stage<BW_BA>(b, GLRGTV_SLOT_BWD_BA, stream))) return rc;
        w.z = x; w.src0 = gbA;
        if ((rc = gw_stage<BW_BA>(w, GLRGTV_SLOT_GW, stream))) return rc;
        glrgtv_shape sc = *s;
        sc.H /= 2; sc.W /= 2;
        GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_WEIGHTS, stream);
        if ((rc = glr_block_weights_bwd(s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, w.gwT0, w.gwL0, gfeat0,
                                        gr->gtv0_M, gr->glr0_M, stream))) return rc;
        rc = glr_block_weights_bwd(&sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, w.gwT1, w.gwL1, gfeat1,
                                   gr->gtv1_M, gr->glr1_M, stream);
        GLR_PROF_END(GLRGTV_SLOT_BWD_WEIGHTS, stream);
        return rc;
    }
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

    // edge weights -> features: one launch per resolution covers both operator families
    glrgtv_shape sc = *s;
    sc.H /= 2; sc.W /= 2;
    (void)o_scr;
    GLR_PROF_BEGIN(GLRGTV_SLOT_BWD_WEIGHTS, stream);
    if ((rc = glr_block_weights_bwd(s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, a.gwT0, a.gwL0, gfeat0,
                                    gr->gtv0_M, gr->glr0_M, stream))) return rc;
    rc = glr_block_weights_bwd(&sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, a.gwT1, a.gwL1, gfeat1,
                               gr->gtv1_M, gr->glr1_M, stream);
    GLR_PROF_END(GLRGTV_SLOT_BWD_WEIGHTS, stream);
    return rc;
}
