// tile.cuh - shared-memory planes and quad-wide stencil stages of the fused block kernels (3x3 cross).
//
// A CTA owns one (batch, graph) pair and one TH x TW tile of the fine grid and walks the graph's F signal
// channels.  Every intermediate of the operator chain lives in a shared-memory PLANE: `rows` x PITCH floats
// covering GLOBAL rows [h0 - R, h0 + TH + R) (R = the plane's halo) and GLOBAL columns [w0 - 8, w0 - 8 + PITCH)
// - every plane of one resolution shares the same column origin and pitch (TW + 16), so that columns line
// up across planes and every 4-column QUAD (local column multiple of 4) is a 16-byte aligned float4.
// Stages work a quad at a time: one thread produces 4 horizontally adjacent outputs from float4 loads of
// the centre / upper / lower quads plus two scalars, which is what keeps the shared-memory instruction
// count (the bottleneck of a 5-point stencil) at ~1/3 of the FMA count.
//
// Pixels outside the image Omega follow two conventions:
//   * clamp-extended planes hold X[cl(p)]   (what replicate-padded gathers read: inputs of S, L, C, P);
//   * zero-extended  planes hold 0          (what the transposed stencils St, Ct read).
// A quad whose four centres are inside Omega takes the fast path; any other quad falls back to a
// per-element path that evaluates the stencil at the CLAMPED centre (clamp-extended result) or writes 0.
//
// Edge order of the cross window (V1X0:26-30, 42-49): e=0 U(-1,0), e=1 L(0,-1), e=2 R(0,1), e=3 D(1,0).
#pragma once
#include "common.cuh"

#ifdef GLRGTV_EMU
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

// Item loops.  The CTA size NT is a compile-time constant carried by the geometry type, so trip counts are known
// and the loops unroll: the independent items of one thread interleave, which is where the ILP comes from.
// (Emulation build: one thread walks all items.)
#ifdef GLRGTV_EMU
#define TILE_LOOP_NT(NT, i, n) for (int i = 0; i < (n); ++i)
#else
#define TILE_LOOP_NT(NT, i, n) \
    _Pragma("unroll") for (int k_##i = 0, i = threadIdx.x; k_##i < ((n) + (NT)-1) / (NT); ++k_##i, i += (NT)) if (i < (n))
#endif
#define TILE_LOOP(i, n) TILE_LOOP_NT(G::NT, i, n)
#define COL0 8  // local column of the tile's first pixel

// geometry of one resolution of a tile: image size, tile origin (global), tile size
template <int TR_, int TC_, int NT_>
struct Geo {
    static constexpr int TR = TR_, TC = TC_, P = TC_ + 16, NT = NT_;
    int H, W, h0, w0;
    // quads [q0, q1) cover local columns [COL0 - R, COL0 + TC + R)
    static constexpr int q0(int R) { return (COL0 - R) / 4; }
    static constexpr int q1(int R) { return (COL0 + TC + R + 3) / 4; }
    static constexpr int nq(int R) { return q1(R) - q0(R); }
    static constexpr int rows(int R) { return TR + 2 * R; }
    static constexpr int items(int R) { return rows(R) * nq(R); }
    static constexpr int floats(int R) { return rows(R) * P; }
    // global coordinates of local (row r of a halo-R plane, column c)
    __device__ __forceinline__ int gh(int r, int R) const { return h0 - R + r; }
    __device__ __forceinline__ int gw(int c) const { return w0 - COL0 + c; }
    __device__ __forceinline__ bool quad_inside(int h, int w) const { return h >= 0 && h < H && w >= 0 && w + 3 < W; }
    __device__ __forceinline__ bool inside(int h, int w) const { return h >= 0 && h < H && w >= 0 && w < W; }
};

// a plane with halo R of geometry G: element (global h, w) lives at p[(h - h0 + R) * P + (w - w0 + COL0)]
template <class G, int R_>
struct Plane {
    static constexpr int R = R_, P = G::P;
    float* p;
    __device__ __forceinline__ float* at(const G& g, int h, int w) const { return p + (h - g.h0 + R) * P + (w - g.w0 + COL0); }
    __device__ __forceinline__ float* lrc(int r, int c) const { return p + r * P + c; }  // local row / column
};
template <class G, int R>
__device__ __forceinline__ Plane<G, R> carve(float*& cur) {
    Plane<G, R> pl{cur};
    cur += G::floats(R);
    return pl;
}

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// 5-point neighbourhood of a quad: centre, up, down quads and the two scalars left / right of the centre quad
struct N5 {
    float c[4], u[4], d[4], l, r;
    // value of the left / right neighbour of element j
    __device__ __forceinline__ float L(int j) const { return j == 0 ? l : c[j - 1]; }
    __device__ __forceinline__ float Rr(int j) const { return j == 3 ? r : c[j + 1]; }
};
template <int P>
__device__ __forceinline__ void ld_n5(const float* s, N5& n) {
    ld4(s, n.c); ld4(s - P, n.u); ld4(s + P, n.d);
    n.l = s[-1]; n.r = s[4];
}

// decompose a loop index into (row, quad) of a halo-R region
#define QUAD_ITEM(G, R, i, r, c)                      \
    const int r = (i) / G::nq(R);                     \
    const int c = 4 * (G::q0(R) + (i) % G::nq(R))

// ------------------------------------------------------------------ loads from global memory
// plane[h,w] = src[cl(h), cl(w)]  (CLAMP) or src[h,w] inside / 0 outside (zero-extended)
template <bool CLAMP, class G, int R>
__device__ __forceinline__ void load_plane(const G& g, const Plane<G, R>& dst, const float* __restrict__ src) {
    const bool vec = (g.W & 3) == 0;
    TILE_LOOP(i, G::items(R)) {
        QUAD_ITEM(G, R, i, r, c);
        const int h = g.gh(r, R), w = g.gw(c);
        float v[4];
        if (vec && g.quad_inside(h, w)) {
            float4 t = *reinterpret_cast<const float4*>(src + (size_t)h * g.W + w);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CLAMP) v[j] = src[(size_t)glr_clampi(h, 0, g.H - 1) * g.W + glr_clampi(w + j, 0, g.W - 1)];
                else v[j] = g.inside(h, w + j) ? src[(size_t)h * g.W + w + j] : 0.f;
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// ------------------------------------------------------------------ forward stages
__device__ __forceinline__ float s_elem(const float* c, int P, const StatsTaps k) {
    return k.kc * c[0] + k.kr * c[1] + k.kd * c[P] + k.ku * c[-P] + k.kl * c[-1];
}
// S with two tap sets from one read: clamp-extended src (halo RS) -> clamp-extended dA, dB (halo RD)
template <bool TWO, class G, int RD, int RS>
__device__ __forceinline__ void stage_S(const G& g, const Plane<G, RD>& dA, const StatsTaps kA, const Plane<G, RD>& dB,
                                        const StatsTaps kB, const Plane<G, RS>& src) {
    static_assert(RS >= RD + 1, "halo");
    TILE_LOOP(i, G::items(RD)) {
        QUAD_ITEM(G, RD, i, r, c);
        const int h = g.gh(r, RD), w = g.gw(c);
        float a[4], b[4];
        if (g.quad_inside(h, w)) {
            N5 n;
            ld_n5<G::P>(src.lrc(r + RS - RD, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[j] = kA.kc * n.c[j] + kA.kr * n.Rr(j) + kA.kd * n.d[j] + kA.ku * n.u[j] + kA.kl * n.L(j);
                if (TWO) b[j] = kB.kc * n.c[j] + kB.kr * n.Rr(j) + kB.kd * n.d[j] + kB.ku * n.u[j] + kB.kl * n.L(j);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* q = src.at(g, glr_clampi(h, 0, g.H - 1), glr_clampi(w + j, 0, g.W - 1));
                a[j] = s_elem(q, G::P, kA);
                if (TWO) b[j] = s_elem(q, G::P, kB);
            }
        }
        st4(dA.lrc(r, c), a);
        if (TWO) st4(dB.lrc(r, c), b);
    }
}

// 2x2 mean: fine clamp-extended src (halo 2*RD) -> coarse clamp-extended dst (halo RD)
template <class GC, class GF, int RD>
__device__ __forceinline__ void stage_pool(const GC& gc, const Plane<GC, RD>& dst, const Plane<GF, 2 * RD>& src) {
    TILE_LOOP_NT(GC::NT, i, GC::items(RD)) {
        QUAD_ITEM(GC, RD, i, r, c);
        const int h = gc.gh(r, RD), w = gc.gw(c);
        float v[4];
        if (gc.quad_inside(h, w)) {
            const float* q = src.lrc(2 * r, 2 * c - COL0);
            float a0[4], a1[4], b0[4], b1[4];
            ld4(q, a0); ld4(q + 4, a1); ld4(q + GF::P, b0); ld4(q + GF::P + 4, b1);
            v[0] = 0.25f * (a0[0] + a0[1] + b0[0] + b0[1]);
            v[1] = 0.25f * (a0[2] + a0[3] + b0[2] + b0[3]);
            v[2] = 0.25f * (a1[0] + a1[1] + b1[0] + b1[1]);
            v[3] = 0.25f * (a1[2] + a1[3] + b1[2] + b1[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int hc = glr_clampi(h, 0, gc.H - 1), wc = glr_clampi(w + j, 0, gc.W - 1);
                const float* q = src.lrc(2 * (hc - gc.h0 + RD), 2 * (wc - gc.w0 + COL0) - COL0);
                v[j] = 0.25f * (q[0] + q[1] + q[GF::P] + q[GF::P + 1]);
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// four weight planes of one graph (same halo), zero-extended
template <class G, int R>
struct WPl {
    Plane<G, R> e[4];
};
template <class G, int R>
__device__ __forceinline__ WPl<G, R> carve_w(float*& cur) {
    WPl<G, R> w;
#pragma unroll
    for (int e = 0; e < 4; ++e) w.e[e] = carve<G, R>(cur);
    return w;
}
template <class G, int R>
__device__ __forceinline__ void load_weights(const G& g, const WPl<G, R>& w, const float* __restrict__ wsrc) {
#pragma unroll
    for (int e = 0; e < 4; ++e) load_plane<false>(g, w.e[e], wsrc + (size_t)e * g.H * g.W);
}

// L: s clamp-extended (halo RS) -> dst zero-extended (halo RD): s - sum_e w_e s[n_e]; weights halo RW >= RD
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void stage_L(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& s, const WPl<G, RW>& w) {
    TILE_LOOP(i, G::items(RD)) {
        QUAD_ITEM(G, RD, i, r, c);
        const int h = g.gh(r, RD), x = g.gw(c);
        float v[4];
        if (g.quad_inside(h, x)) {
            N5 n;
            ld_n5<G::P>(s.lrc(r + RS - RD, c), n);
            float w0[4], w1[4], w2[4], w3[4];
            ld4(w.e[0].lrc(r + RW - RD, c), w0); ld4(w.e[1].lrc(r + RW - RD, c), w1);
            ld4(w.e[2].lrc(r + RW - RD, c), w2); ld4(w.e[3].lrc(r + RW - RD, c), w3);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = n.c[j] - (w0[j] * n.u[j] + w1[j] * n.L(j) + w2[j] * n.Rr(j) + w3[j] * n.d[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = 0.f;
                if (g.inside(h, x + j)) {
                    const float* q = s.at(g, h, x + j);
                    v[j] = q[0] - (*w.e[0].at(g, h, x + j) * q[-G::P] + *w.e[1].at(g, h, x + j) * q[-1] +
                                   *w.e[2].at(g, h, x + j) * q[1] + *w.e[3].at(g, h, x + j) * q[G::P]);
                }
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// phi(t) = 2*soft(t,G) - t = t - 2*clamp(t,-G,G)   (V1X0:765-777: epsilon - bias);  phi'(t) = |t|>G ? 1 : -1
__device__ __forceinline__ float glr_phi(float t, float G) { return t - 2.f * fminf(fmaxf(t, -G), G); }
__device__ __forceinline__ float glr_dphi(float t, float G) { return fabsf(t) > G ? 1.f : -1.f; }

// symmetric GTV coefficients: cR[q] = wR[q]^2 + wL[q+(0,1)]^2, cD[q] = wD[q]^2 + wU[q+(1,0)]^2 (0 outside Omega),
// computed straight from global memory into two planes of halo R.
template <class G, int R>
__device__ __forceinline__ void load_gtv_coeffs(const G& g, const Plane<G, R>& cR, const Plane<G, R>& cD,
                                                const float* __restrict__ wsrc) {
    const size_t HW = (size_t)g.H * g.W;
    TILE_LOOP(i, G::items(R) * 4) {
        const int q = i >> 2, j = i & 3;
        QUAD_ITEM(G, R, q, r, c);
        const int h = g.gh(r, R), w = g.gw(c) + j;
        float vr = 0.f, vd = 0.f;
        if (g.inside(h, w)) {
            const size_t o = (size_t)h * g.W + w;
            const float wr = wsrc[2 * HW + o], wd = wsrc[3 * HW + o];
            const float wl = w + 1 < g.W ? wsrc[1 * HW + o + 1] : 0.f;
            const float wu = h + 1 < g.H ? wsrc[0 * HW + o + g.W] : 0.f;
            vr = wr * wr + wl * wl;
            vd = wd * wd + wu * wu;
        }
        *cR.lrc(r, c + j) = vr;
        *cD.lrc(r, c + j) = vd;
    }
}

// linear GTV core  o = Ct C s = sum_n c_n (s[q]-s[n])  (self-adjoint): s clamp-extended (halo RS),
// coefficient planes halo RC, dst zero-extended (halo RD).  (SURVEY B.5/B.6 combined.)
template <class G, int RD, int RS, int RC>
__device__ __forceinline__ void stage_gtv_lin(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& s,
                                              const Plane<G, RC>& cR, const Plane<G, RC>& cD) {
    static_assert(RC >= RD + 1 && RS >= RD + 1, "halo");
    TILE_LOOP(i, G::items(RD)) {
        QUAD_ITEM(G, RD, i, r, c);
        const int h = g.gh(r, RD), x = g.gw(c);
        float v[4];
        if (g.quad_inside(h, x)) {
            N5 n;
            ld_n5<G::P>(s.lrc(r + RS - RD, c), n);
            float cr[4], cd[4], cu[4];
            const float* pr = cR.lrc(r + RC - RD, c);
            ld4(pr, cr); ld4(cD.lrc(r + RC - RD, c), cd); ld4(cD.lrc(r + RC - RD - 1, c), cu);
            const float crl = pr[-1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float cl_ = j == 0 ? crl : cr[j - 1];
                v[j] = cr[j] * (n.c[j] - n.Rr(j)) + cl_ * (n.c[j] - n.L(j)) + cd[j] * (n.c[j] - n.d[j]) + cu[j] * (n.c[j] - n.u[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = 0.f;
                if (g.inside(h, x + j)) {
                    const float* q = s.at(g, h, x + j);
                    const float* a = cR.at(g, h, x + j);
                    const float* b = cD.at(g, h, x + j);
                    v[j] = a[0] * (q[0] - q[1]) + a[-1] * (q[0] - q[-1]) + b[0] * (q[0] - q[G::P]) + b[-G::P] * (q[0] - q[-G::P]);
                }
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// thresholded GTV core with raw weights (zero-extended, halo RW >= RD+1):
//   o[q] = sum_n [ wa phi(wa d) + wb phi(wb d) ],  d = s[q]-s[n], wa = w_{q->n}[q], wb = w_{n->q}[n]
__device__ __forceinline__ float gtv_thr_elem(const float* q, int P, const float* w0, const float* w1, const float* w2,
                                              const float* w3, float G) {
    const float v = q[0];
    float d, r;
    d = v - q[-P]; r = w0[0] * glr_phi(w0[0] * d, G) + w3[-P] * glr_phi(w3[-P] * d, G);
    d = v - q[-1]; r += w1[0] * glr_phi(w1[0] * d, G) + w2[-1] * glr_phi(w2[-1] * d, G);
    d = v - q[1];  r += w2[0] * glr_phi(w2[0] * d, G) + w1[1] * glr_phi(w1[1] * d, G);
    d = v - q[P];  r += w3[0] * glr_phi(w3[0] * d, G) + w0[P] * glr_phi(w0[P] * d, G);
    return r;
}
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void stage_gtv_thr(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& s,
                                              const WPl<G, RW>& w, float Gam) {
    static_assert(RW >= RD + 1 && RS >= RD + 1, "halo");
    TILE_LOOP(i, G::items(RD) * 4) {   // element-wise: the thresholded core is ALU-heavy, not load-heavy
        const int qd = i >> 2, j = i & 3;
        QUAD_ITEM(G, RD, qd, r, c);
        const int h = g.gh(r, RD), x = g.gw(c) + j;
        float v = 0.f;
        if (g.inside(h, x))
            v = gtv_thr_elem(s.at(g, h, x), G::P, w.e[0].at(g, h, x), w.e[1].at(g, h, x), w.e[2].at(g, h, x),
                             w.e[3].at(g, h, x), Gam);
        *dst.lrc(r, c + j) = v;
    }
}
// same core with raw weights and phi = identity (used where only the raw planes are resident)
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void stage_gtv_lin_raw(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& s,
                                                  const WPl<G, RW>& w) {
    TILE_LOOP(i, G::items(RD) * 4) {
        const int qd = i >> 2, j = i & 3;
        QUAD_ITEM(G, RD, qd, r, c);
        const int h = g.gh(r, RD), x = g.gw(c) + j;
        float v = 0.f;
        if (g.inside(h, x)) {
            const float* q = s.at(g, h, x);
            const float *w0 = w.e[0].at(g, h, x), *w1 = w.e[1].at(g, h, x), *w2 = w.e[2].at(g, h, x), *w3 = w.e[3].at(g, h, x);
            constexpr int P = G::P;
            v = (w0[0] * w0[0] + w3[-P] * w3[-P]) * (q[0] - q[-P]) + (w1[0] * w1[0] + w2[-1] * w2[-1]) * (q[0] - q[-1]) +
                (w2[0] * w2[0] + w1[1] * w1[1]) * (q[0] - q[1]) + (w3[0] * w3[0] + w0[P] * w0[P]) * (q[0] - q[P]);
        }
        *dst.lrc(r, c + j) = v;
    }
}

// St at a quad from a zero-extended plane: sum_t k_t y[q - o_t]
template <int P>
__device__ __forceinline__ void St_quad(const float* y, const StatsTaps k, float (&v)[4]) {
    N5 n;
    ld_n5<P>(y, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = k.kc * n.c[j] + k.kr * n.L(j) + k.kd * n.u[j] + k.ku * n.d[j] + k.kl * n.Rr(j);
}
__device__ __forceinline__ float St_elem(const float* c, int P, const StatsTaps k) {
    return k.kc * c[0] + k.kr * c[-1] + k.kd * c[-P] + k.ku * c[P] + k.kl * c[1];
}

// ------------------------------------------------------------------ adjoint (VJP) stages, SURVEY B.2-B.6
// For a unit offset d the adjoint of the clamped gather x[cl(p+d)] is
//   sum_{p: cl(p+d)=q} v[p] = v_zero[q-d] + [q+d outside] v[q]     (the border pixel also collects the
// tap that was replicated onto it).

// VJP of St wrt its input = "S with zero padding": out[p] = scale * sum_t k_t g[p + o_t], g zero-extended.
// CLAMP_OUT: clamp-extended result (feeds the GTV core), else zero-extended (feeds the L adjoint).
template <bool CLAMP_OUT, class G, int RD, int RS>
__device__ __forceinline__ void stage_Szero(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& src,
                                            const StatsTaps k, float scale) {
    static_assert(RS >= RD + 1, "halo");
    TILE_LOOP(i, G::items(RD)) {
        QUAD_ITEM(G, RD, i, r, c);
        const int h = g.gh(r, RD), w = g.gw(c);
        float v[4];
        if (g.quad_inside(h, w)) {
            N5 n;
            ld_n5<G::P>(src.lrc(r + RS - RD, c), n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = scale * (k.kc * n.c[j] + k.kr * n.Rr(j) + k.kd * n.d[j] + k.ku * n.u[j] + k.kl * n.L(j));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CLAMP_OUT) {
                    v[j] = scale * s_elem(src.at(g, glr_clampi(h, 0, g.H - 1), glr_clampi(w + j, 0, g.W - 1)), G::P, k);
                } else {
                    v[j] = g.inside(h, w + j) ? scale * s_elem(src.at(g, h, w + j), G::P, k) : 0.f;
                }
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// VJP of L wrt its input, zero-extended: gs[q] = gl[q] - sum_n w_{n->q}[n] gl[n] - sum_{e: q+d_e outside} w_e[q] gl[q]
// gl zero-extended (halo RS >= RD+1), w zero-extended (halo RW >= RD+1).
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void stage_L_adj(const G& g, const Plane<G, RD>& dst, const Plane<G, RS>& gl, const WPl<G, RW>& w) {
    static_assert(RS >= RD + 1 && RW >= RD + 1, "halo");
    TILE_LOOP(i, G::items(RD)) {
        QUAD_ITEM(G, RD, i, r, c);
        const int h = g.gh(r, RD), x = g.gw(c);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = 0.f;
        if (g.quad_inside(h, x)) {
            N5 n;
            ld_n5<G::P>(gl.lrc(r + RS - RD, c), n);
            const int rw = r + RW - RD;
            float wu[4], wd[4], wl[4], wr[4];
            ld4(w.e[3].lrc(rw - 1, c), wu);   // edge D of the upper neighbour points at q
            ld4(w.e[0].lrc(rw + 1, c), wd);   // edge U of the lower neighbour
            const float* pR = w.e[2].lrc(rw, c);  // edge R of the left neighbour: columns c-1 .. c+2
            const float* pL = w.e[1].lrc(rw, c);  // edge L of the right neighbour: columns c+1 .. c+4
            ld4(pR, wr); ld4(pL, wl);
            const float wr_m1 = pR[-1], wl_p4 = pL[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wrn = j == 0 ? wr_m1 : wr[j - 1], wln = j == 3 ? wl_p4 : wl[j + 1];
                v[j] = n.c[j] - (wu[j] * n.u[j] + wrn * n.L(j) + wln * n.Rr(j) + wd[j] * n.d[j]);
            }
            if (h == 0 || h == g.H - 1 || x == 0 || x + 3 == g.W - 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float self = 0.f;
                    if (h == 0) self += *w.e[0].lrc(rw, c + j);
                    if (x + j == 0) self += *w.e[1].lrc(rw, c + j);
                    if (x + j == g.W - 1) self += *w.e[2].lrc(rw, c + j);
                    if (h == g.H - 1) self += *w.e[3].lrc(rw, c + j);
                    v[j] -= self * n.c[j];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = x + j;
                if (!g.inside(h, xx)) continue;
                const float* q = gl.at(g, h, xx);
                constexpr int P = G::P;
                float rr = q[0] - (w.e[3].at(g, h, xx)[-P] * q[-P] + w.e[2].at(g, h, xx)[-1] * q[-1] +
                                   w.e[1].at(g, h, xx)[1] * q[1] + w.e[0].at(g, h, xx)[P] * q[P]);
                float self = 0.f;
                if (h == 0) self += *w.e[0].at(g, h, xx);
                if (xx == 0) self += *w.e[1].at(g, h, xx);
                if (xx == g.W - 1) self += *w.e[2].at(g, h, xx);
                if (h == g.H - 1) self += *w.e[3].at(g, h, xx);
                v[j] = rr - self * q[0];
            }
        }
        st4(dst.lrc(r, c), v);
    }
}

// VJP of S wrt its input at one pixel / one quad, from a zero-extended gs plane
__device__ __forceinline__ float S_adj_elem(const float* c, int P, const StatsTaps k, int h, int w, int H, int W) {
    const float v = c[0];
    float r = k.kc * v + k.kr * c[-1] + k.kd * c[-P] + k.ku * c[P] + k.kl * c[1];
    float self = 0.f;
    if (w == W - 1) self += k.kr;
    if (h == H - 1) self += k.kd;
    if (h == 0) self += k.ku;
    if (w == 0) self += k.kl;
    return r + self * v;
}
template <int P>
__device__ __forceinline__ void S_adj_quad(const float* gs, const StatsTaps k, int h, int w, int H, int W, float (&v)[4]) {
    N5 n;
    ld_n5<P>(gs, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float self = 0.f;
        if (w + j == W - 1) self += k.kr;
        if (h == H - 1) self += k.kd;
        if (h == 0) self += k.ku;
        if (w + j == 0) self += k.kl;
        v[j] = (k.kc + self) * n.c[j] + k.kr * n.L(j) + k.kd * n.u[j] + k.ku * n.d[j] + k.kl * n.Rr(j);
    }
}

// VJP of the thresholded GTV core wrt s, ADDED into dst (which already holds the linear part):
//   gs[q] += sum_n (go[q]-go[n]) (wa^2 phi'(wa d) + wb^2 phi'(wb d)),  d = s[q]-s[n]
// go, s clamp-extended, w zero-extended; all halos >= RD+1.
template <class G, int RD, int RG, int RS, int RW>
__device__ __forceinline__ void stage_gtv_thr_adj_add(const G& g, const Plane<G, RD>& dst, const Plane<G, RG>& go,
                                                      const Plane<G, RS>& s, const WPl<G, RW>& w, float Gam) {
    TILE_LOOP(i, G::items(RD) * 4) {
        const int qd = i >> 2, j = i & 3;
        QUAD_ITEM(G, RD, qd, r, c);
        const int h = g.gh(r, RD), x = g.gw(c) + j;
        if (!g.inside(h, x)) continue;
        constexpr int P = G::P;
        const float* cg = go.at(g, h, x);
        const float* cs = s.at(g, h, x);
        const float *w0 = w.e[0].at(g, h, x), *w1 = w.e[1].at(g, h, x), *w2 = w.e[2].at(g, h, x), *w3 = w.e[3].at(g, h, x);
        const float gv = cg[0], sv = cs[0];
        float d, rr;
        d = sv - cs[-P]; rr = (gv - cg[-P]) * (w0[0] * w0[0] * glr_dphi(w0[0] * d, Gam) + w3[-P] * w3[-P] * glr_dphi(w3[-P] * d, Gam));
        d = sv - cs[-1]; rr += (gv - cg[-1]) * (w1[0] * w1[0] * glr_dphi(w1[0] * d, Gam) + w2[-1] * w2[-1] * glr_dphi(w2[-1] * d, Gam));
        d = sv - cs[1];  rr += (gv - cg[1]) * (w2[0] * w2[0] * glr_dphi(w2[0] * d, Gam) + w1[1] * w1[1] * glr_dphi(w1[1] * d, Gam));
        d = sv - cs[P];  rr += (gv - cg[P]) * (w3[0] * w3[0] * glr_dphi(w3[0] * d, Gam) + w0[P] * w0[P] * glr_dphi(w0[P] * d, Gam));
        *dst.lrc(r, c + j) += rr;
    }
}

// block-wide sums of N values at once; results valid in thread 0.  red: >= 32*N floats.
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float* red) {
#ifndef GLRGTV_EMU
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < N; ++k) red[wid * N + k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < N; ++k) {
            float s = 0.f;
            for (int j = 0; j < nw; ++j) s += red[j * N + k];
            v[k] = s;
        }
#endif
}

template <class G, int R>
__device__ __forceinline__ Plane<G, R> plane_at(float* smem, int off) { return Plane<G, R>{smem + off}; }
template <class G, int R>
__device__ __forceinline__ WPl<G, R> wplanes_at(float* smem, int off) {
    WPl<G, R> w;
#pragma unroll
    for (int e = 0; e < 4; ++e) w.e[e].p = smem + off + e * G::floats(R);
    return w;
}
static inline int glr_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// zero-extended 2x2 mean of a zero-extended fine plane (VJP of P^T is P): coarse halo RD from fine halo 2*RD
template <class GC, class GF, int RD>
__device__ __forceinline__ void stage_pool_zero(const GC& gc, const Plane<GC, RD>& dst, const Plane<GF, 2 * RD>& src) {
    TILE_LOOP_NT(GC::NT, i, GC::items(RD)) {
        QUAD_ITEM(GC, RD, i, r, c);
        const int h = gc.gh(r, RD), w = gc.gw(c);
        float v[4];
        const float* q = src.lrc(2 * r, 2 * c - COL0);
        float a0[4], a1[4], b0[4], b1[4];
        ld4(q, a0); ld4(q + 4, a1); ld4(q + GF::P, b0); ld4(q + GF::P + 4, b1);
        v[0] = 0.25f * (a0[0] + a0[1] + b0[0] + b0[1]);
        v[1] = 0.25f * (a0[2] + a0[3] + b0[2] + b0[3]);
        v[2] = 0.25f * (a1[0] + a1[1] + b1[0] + b1[1]);
        v[3] = 0.25f * (a1[2] + a1[3] + b1[2] + b1[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (!gc.inside(h, w + j)) v[j] = 0.f;
        st4(dst.lrc(r, c), v);
    }
}
