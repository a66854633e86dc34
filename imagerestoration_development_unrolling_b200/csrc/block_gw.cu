// block_gw.cu - edge-weight gradients of one backward stage of the fused block.
//
// The weight gradients of a stage need only the FIRST-level stencils of its two operands,
//     s_K = S_K z (clamp-extended),   h_K = a_K S0_K g   (g = the stage's upstream gradient),
//   L :      gw_e[p] -= sum_f  h_L[p] * s_L[n_e(p)]
//   T lin :  gw_e[p] += sum_f  2 w_e (h_T[p] - h_T[n_e]) (s_T[p] - s_T[n_e])
//   T thr :  gw_e[p] += sum_f  D phi(w d) + D w phi'(w d) d,   D = h_T[p] - h_T[n_e], d = s_T[p] - s_T[n_e]   (stage X2, upstream gB)
//            ggamma   += sum   -2 sign(w d) D w [|w d| > Gamma]   (times Gamma: the parameter is a log)
// and they are sums over the F channels of a graph.  The streaming stage kernels (block_stream_bwd.cu) give every
// channel its own walkers, so this reduction lives here: a CTA owns one (batch, graph) and one 32 x 32 tile of one
// resolution, walks the graph's channels with the tile's operands in shared memory, and keeps the eight
// gradient values of each pixel of its quad in registers until the end.  The coarse resolution pools its operands
// (2x2 mean) while loading.
#include "tile.cuh"
#include "stream_bwd.cuh"

enum { GW_X3 = BW_X3, GW_X2 = BW_X2A, GW_X1 = BW_X1, GW_BA = BW_BA };   // GW_X2 covers both parts of stage X2

// The kernel uses the tile geometry and the stencil stages of tile.cuh (32 x 32 tile, pitch-40 planes, one aligned float4
// quad per item, LDS.128), one quad of output pixels per thread; the plane width at each resolution must be a multiple
// of 4 (every shape the streaming backward accepts).  MODE GW_X2 does the linear part (upstream gA) and the thresholded
// part (upstream gB) of stage X2 in one pass: both read z = x1 and the same two source tensors.
#define GQ_NT 256
#ifndef GQ_MINB
#define GQ_MINB 3
#endif
template <int MODE, bool COARSE>
__global__ void __launch_bounds__(GQ_NT, GQ_MINB) k_gw_quad(GwArgs a) {
    GLR_SMEM_DECL(smem);
    constexpr bool HAS_L = MODE != GW_BA, X2 = MODE == GW_X2;
    using GG = Geo<32, 32, GQ_NT, true>;
    constexpr int F2 = GG::floats(2), F1 = GG::floats(1);
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int LH = COARSE ? H / 2 : H, LW = COARSE ? W / 2 : W;
    const int tiles_w = (LW + 31) / 32, tiles_h = (LH + 31) / 32;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    GG gg; gg.H = LH; gg.W = LW; gg.h0 = (tile / tiles_w) * 32; gg.w0 = (tile % tiles_w) * 32;
    const size_t HW = (size_t)H * W, LHW = (size_t)LH * LW;
    // planes (4 guard floats in front: quad 0 of the first row reads the scalar left of it)
    int o = 4;
    auto zt = plane_at<GG, 2>(smem, o); o += F2;
    auto gt = plane_at<GG, 2>(smem, o); o += F2;
    auto g2 = plane_at<GG, 2>(smem, o); o += X2 ? F2 : 0;
    auto sA = plane_at<GG, 1>(smem, o); o += HAS_L ? F1 : 0;
    auto sB = plane_at<GG, 1>(smem, o); o += F1;
    auto hL = plane_at<GG, 1>(smem, o); o += HAS_L ? F1 : 0;
    auto hT = plane_at<GG, 1>(smem, o); o += F1;
    auto h2 = plane_at<GG, 1>(smem, o); o += X2 ? F1 : 0;
    float* red = smem + o + 4;

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

    // this thread's output quad, its raw GTV weights (channel-independent) and gradient accumulators
    const int tid = (int)threadIdx.x, qr = tid >> 3, qc = 4 * (1 + (tid & 7));
    const int ph = gg.h0 + qr, pw = gg.gw(qc);
    const bool ok = ph < LH && pw < LW;
    float we[4][4], accT[4][4], accL[4][4], gam = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (ok) ld4(wT + e * LHW + (size_t)ph * LW + pw, we[e]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { if (!ok) we[e][j] = 0.f; accT[e][j] = 0.f; accL[e][j] = 0.f; }
    }
    // one element of an operand at this resolution (coordinates already clamped into the plane)
    auto ld1 = [&](const float* src, int h, int w) -> float {
        if (!COARSE) return src[(size_t)h * W + w];
        const float* q = src + (size_t)(2 * h) * W + 2 * w;
        return 0.25f * (q[0] + q[1] + q[W] + q[W + 1]);
    };
    // one aligned quad at this resolution
    auto ldq = [&](const float* src, int h, int w, float (&v)[4]) {
        if (!COARSE) { ld4(src + (size_t)h * W + w, v); return; }
        const float* q = src + (size_t)(2 * h) * W + 2 * w;
        float t0[4], t1[4], u0[4], u1[4];
        ld4(q, t0); ld4(q + 4, t1); ld4(q + W, u0); ld4(q + W + 4, u1);
        v[0] = 0.25f * (t0[0] + t0[1] + u0[0] + u0[1]); v[1] = 0.25f * (t0[2] + t0[3] + u0[2] + u0[3]);
        v[2] = 0.25f * (t1[0] + t1[1] + u1[0] + u1[1]); v[3] = 0.25f * (t1[2] + t1[3] + u1[2] + u1[3]);
    };

    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const size_t off = ((size_t)b * G * F + c) * HW;
        const StatsTaps kT = glr_load_taps(stT, c), kL = HAS_L ? glr_load_taps(stL, c) : kT;
        // ---- operands on the tile (+) 2: z clamp-extended, the upstream gradient(s) zero outside.  (A register prefetch
        //      of the next channel was measured and does not pay: the phases are short and barrier-bound, not load-bound.)
        for (int i = tid; i < GG::items(2); i += GQ_NT) {
            const Quad q = quad_of<GG, 2>(gg, i);
            float vz[4], v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
            if (q.fast) {
                ldq(a.z + off, q.h, q.w, vz);
                ldq(a.src0 + off, q.h, q.w, v0);
                if (X2) ldq(a.src1 + off, q.h, q.w, v1);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int hc = glr_clampi(q.h, 0, LH - 1), wc = glr_clampi(q.w + j, 0, LW - 1);
                    vz[j] = ld1(a.z + off, hc, wc);
                    if (gg.inside(q.h, q.w + j)) {
                        v0[j] = ld1(a.src0 + off, hc, wc);
                        if (X2) v1[j] = ld1(a.src1 + off, hc, wc);
                    }
                }
            }
            float ga[4], gb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { ga[j] = ca * v0[j] + cb * v1[j]; gb[j] = ca2 * v0[j] + cb2 * v1[j]; }
            st4(zt.lrc(q.r, q.c), vz);
            st4(gt.lrc(q.r, q.c), ga);
            if (X2) st4(g2.lrc(q.r, q.c), gb);
        }
        __syncthreads();
        // ---- first-level stencils on the tile (+) 1
        for (int i = tid; i < GG::items(1); i += GQ_NT) {
            const Quad q = quad_of<GG, 1>(gg, i);
            q_S<HAS_L>(gg, q, HAS_L ? sA : sB, HAS_L ? kL : kT, sB, kT, zt);
            q_Szero<HAS_L, true>(gg, q, hL, kL, aL, hT, kT, aT, gt);
            if (X2) q_Szero<false, true>(gg, q, h2, kT, 0.f, h2, kT, aT, g2);
        }
        __syncthreads();
        // ---- gradients of this thread's quad
        if (ok) {
            N5 ns, nh;
            ld_n5<GG::P>(sB.lrc(qr + 1, qc), ns);
            ld_n5<GG::P>(hT.lrc(qr + 1, qc), nh);
            N5 n2;
            if (X2) ld_n5<GG::P>(h2.lrc(qr + 1, qc), n2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d[4] = {ns.c[j] - ns.u[j], ns.c[j] - ns.L(j), ns.c[j] - ns.Rr(j), ns.c[j] - ns.d[j]};
                const float D[4] = {nh.c[j] - nh.u[j], nh.c[j] - nh.L(j), nh.c[j] - nh.Rr(j), nh.c[j] - nh.d[j]};
#pragma unroll
                for (int e = 0; e < 4; ++e) accT[e][j] += 2.f * we[e][j] * D[e] * d[e];
                if (X2) {
                    const float D2[4] = {n2.c[j] - n2.u[j], n2.c[j] - n2.L(j), n2.c[j] - n2.Rr(j), n2.c[j] - n2.d[j]};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float w = we[e][j], t = w * d[e];
                        accT[e][j] += D2[e] * glr_phi(t, Gam) + D2[e] * w * glr_dphi(t, Gam) * d[e];
                        if (fabsf(t) > Gam) gam += D2[e] * w * (t > 0.f ? -2.f : 2.f);
                    }
                }
            }
            if (HAS_L) {
                N5 na;
                ld_n5<GG::P>(sA.lrc(qr + 1, qc), na);
                float hl[4];
                ld4(hL.lrc(qr + 1, qc), hl);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    accL[0][j] -= hl[j] * na.u[j];
                    accL[1][j] -= hl[j] * na.L(j);
                    accL[2][j] -= hl[j] * na.Rr(j);
                    accL[3][j] -= hl[j] * na.d[j];
                }
            }
        }
        __syncthreads();
    }
    if (ok) {
        const size_t po = (size_t)ph * LW + pw;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!a.assign) {
                float v[4];
                ld4(gwT + e * LHW + po, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) accT[e][j] += v[j];
                if (HAS_L) {
                    ld4(gwL + e * LHW + po, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) accL[e][j] += v[j];
                }
            }
            st4(gwT + e * LHW + po, accT[e]);
            if (HAS_L) st4(gwL + e * LHW + po, accL[e]);
        }
    }
    if (X2) {
        const float tot = block_sum(gam, red);
        if (threadIdx.x == 0 && tot != 0.f) atomicAdd((COARSE ? a.ggamma1 : a.ggamma0) + g, tot * Gam);
    }
}

extern unsigned long long g_glr_stream_launches;
template <int MODE>
int glr_gw_stage(const GwArgs& a, int slot, void* stream) {
    const glrgtv_shape& s = a.s;
    constexpr bool HAS_L = MODE != GW_BA, X2 = MODE == GW_X2;
    using GG = Geo<32, 32, GQ_NT, true>;
    const size_t smem_q = ((X2 ? 3 : 2) * GG::floats(2) + (2 + (HAS_L ? 2 : 0) + (X2 ? 1 : 0)) * GG::floats(1) + 4 + 4 + 32) * sizeof(float);
    for (int lvl = 0; lvl < 2; ++lvl) {
        const int LH = lvl ? s.H / 2 : s.H, LW = lvl ? s.W / 2 : s.W;
        if (LW & 3) return GLRGTV_ERR_UNSUPPORTED;
        const int th = 32, tw = 32;
        const long blocks = (long)((LW + tw - 1) / tw) * ((LH + th - 1) / th) * s.B * s.G;
        if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
        ++g_glr_stream_launches;
        GLR_PROF_BEGIN(slot, stream);
        {
#ifndef GLRGTV_EMU
            static size_t optin0[GLR_MAX_DEVICES] = {0}, optin1[GLR_MAX_DEVICES] = {0};
            if (int rc_ = lvl ? glr_smem_optin(k_gw_quad<MODE, true>, smem_q, optin1) : glr_smem_optin(k_gw_quad<MODE, false>, smem_q, optin0)) return rc_;
#endif
            if (lvl) GLR_LAUNCH_FIBERS((k_gw_quad<MODE, true>), dim3((unsigned)blocks), GQ_NT, smem_q, stream, a);
            else GLR_LAUNCH_FIBERS((k_gw_quad<MODE, false>), dim3((unsigned)blocks), GQ_NT, smem_q, stream, a);
        }
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
