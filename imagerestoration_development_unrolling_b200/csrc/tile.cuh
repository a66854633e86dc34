// tile.cuh - shared-memory planes and quad-wide stencil stages of the fused block kernels (3x3 cross).
//
// A CTA owns one (batch, graph) pair and one TR x TC tile of the fine grid and walks the graph's F signal
// channels.  Every intermediate of the operator chain lives in a shared-memory PLANE with halo R: rows
// [h0-R, h0+TR+R) x columns [w0-4, w0+TC+4), row-major with pitch P = TC+8 (40 fine / 24 coarse).  All planes
// of one resolution share the column origin, so columns line up across planes and every 4-column QUAD (local
// column multiple of 4) is one aligned float4.  Column halos go up to 3 (columns -4 and TC+3 are slack).
//
// Why these numbers: a stage item is one quad; a warp's 32 lanes walk consecutive (row, quad) items, NQ = P/4
// quads per row.  A float4 shared load is conflict-free when the 8 lanes of a quarter-warp hit 8 distinct
// 16-byte bank groups, i.e. when the bank group is a function of the item index mod 8, which holds iff
// (P/4) == NQ (mod 8) - true for P=40,NQ=10 and P=24,NQ=6 (and for the 8-quad epilogue rows trivially).
//
// Pixels outside the image Omega follow two conventions:
//   * clamp-extended planes hold X[cl(p)]   (what replicate-padded gathers read: inputs of S, L, C, P);
//   * zero-extended  planes hold 0          (what the transposed stencils St, Ct read).
// A quad whose four centres are inside Omega takes the fast path; any other quad falls back to a per-element
// path that evaluates the stencil at the CLAMPED centre (clamp-extended result) or writes 0.
//
// Edge order of the cross window (V1X0:26-30, 42-49): e=0 U(-1,0), e=1 L(0,-1), e=2 R(0,1), e=3 D(1,0).
#pragma once
#include "common.cuh"

#ifdef GLRGTV_EMU
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

// Item loops.  The CTA size NT is a compile-time constant, so trip counts are known and the loops unroll: the
// independent items of one thread interleave, which is where the ILP comes from.  (Emulation: one thread walks
// all items.)
#ifdef GLRGTV_EMU
#define TILE_LOOP_NT(NT, i, n) for (int i = 0; i < (n); ++i)
#else
#define TILE_LOOP_NT(NT, i, n) \
    _Pragma("unroll") for (int k_##i = 0, i = threadIdx.x; k_##i < ((n) + (NT)-1) / (NT); ++k_##i, i += (NT)) if (i < (n))
#endif
// same loop with the items dealt out from the LAST thread downwards: a phase that runs a fine loop and then a
// coarse loop gives the extra coarse items to the threads the fine loop left with one item less
#ifdef GLRGTV_EMU
#define TILE_LOOP_REV(NT, i, n) for (int i = 0; i < (n); ++i)
#else
#define TILE_LOOP_REV(NT, i, n) \
    _Pragma("unroll") for (int k_##i = 0, i = (NT)-1 - (int)threadIdx.x; k_##i < ((n) + (NT)-1) / (NT); ++k_##i, i += (NT)) if (i < (n))
#endif
#define COL0 4  // local column of the tile's first pixel

// geometry of one resolution of a tile: image size, tile origin (global), tile size
// GEN_ = true: generic borders (any even W); false: W is a multiple of 4 at this resolution, so every quad lies fully
// inside or fully outside the image and the stage bodies are branch-free (see "branch-free borders" below)
template <int TR_, int TC_, int NT_, bool GEN_ = true>
struct Geo {
    static constexpr int TR = TR_, TC = TC_, P = TC_ + 8, NT = NT_, NQ = P / 4;
    static constexpr bool GEN = GEN_;
    int H, W, h0, w0;
    static constexpr int rows(int R) { return TR + 2 * R; }
    static constexpr int items(int R) { return rows(R) * NQ; }   // quads of a halo-R plane
    static constexpr int floats(int R) { return rows(R) * P; }
    __device__ __forceinline__ int gh(int r, int R) const { return h0 - R + r; }   // global row of plane row r
    __device__ __forceinline__ int gw(int c) const { return w0 - COL0 + c; }       // global column of local column c
    __device__ __forceinline__ bool quad_inside(int h, int w) const { return h >= 0 && h < H && w >= 0 && w + 3 < W; }
    __device__ __forceinline__ bool inside(int h, int w) const { return h >= 0 && h < H && w >= 0 && w < W; }
};

// a plane with halo R: element (global h, w) lives at p[(h - h0 + R) * P + (w - w0 + COL0)]
template <class G, int R_>
struct Plane {
    static constexpr int R = R_, P = G::P;
    float* p;
    __device__ __forceinline__ float* at(const G& g, int h, int w) const { return p + (h - g.h0 + R) * P + (w - g.w0 + COL0); }
    __device__ __forceinline__ float* lrc(int r, int c) const { return p + r * P + c; }  // local row / column
};
template <class G, int R>
__device__ __forceinline__ Plane<G, R> plane_at(float* smem, int off) { return Plane<G, R>{smem + off}; }

// four weight planes of one graph (same halo), zero-extended
template <class G, int R>
struct WPl {
    Plane<G, R> e[4];
};
template <class G, int R>
__device__ __forceinline__ WPl<G, R> wplanes_at(float* smem, int off) {
    WPl<G, R> w;
#pragma unroll
    for (int e = 0; e < 4; ++e) w.e[e].p = smem + off + e * G::floats(R);
    return w;
}

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    GLR_CHECK_ALIGN(p, 16);
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    GLR_CHECK_ALIGN(p, 16);
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
static inline int glr_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// 5-point neighbourhood of a quad: centre, up, down quads and the two scalars left / right of the centre quad
struct N5 {
    float c[4], u[4], d[4], l, r;
    __device__ __forceinline__ float L(int j) const { return j == 0 ? l : c[j - 1]; }    // left neighbour of element j
    __device__ __forceinline__ float Rr(int j) const { return j == 3 ? r : c[j + 1]; }   // right neighbour
};
template <int P>
__device__ __forceinline__ void ld_n5(const float* s, N5& n) {
    ld4(s, n.c); ld4(s - P, n.u); ld4(s + P, n.d);
    n.l = s[-1]; n.r = s[4];
}

// one quad of a halo-R plane: local row / column, global coordinates of its first pixel, fast-path flag
struct Quad {
    int r, c, h, w;
    bool fast;     // all four pixels inside the image
    bool border;   // fast and touching the image border (its values are replicated outwards in branch-free mode)
};
template <class G, int R>
__device__ __forceinline__ Quad quad_of(const G& g, int i) {
    Quad q;
    q.r = i / G::NQ;
    q.c = 4 * (i % G::NQ);
    q.h = g.gh(q.r, R);
    q.w = g.gw(q.c);
    q.fast = g.quad_inside(q.h, q.w);
    q.border = q.fast && (q.h == 0 || q.h == g.H - 1 || q.w == 0 || q.w + 4 == g.W);
    return q;
}
// the quads of the tile itself (halo 0): NQ-2 per row, starting at local column 4
template <class G>
__device__ __forceinline__ Quad tile_quad_of(const G& g, int i) {
    Quad q;
    q.r = i / (G::NQ - 2);
    q.c = 4 * (1 + i % (G::NQ - 2));
    q.h = g.h0 + q.r;
    q.w = g.gw(q.c);
    q.fast = g.quad_inside(q.h, q.w);
    q.border = false;
    return q;
}

// ---- branch-free borders (G::GEN == false).  Outside quads are skipped altogether: zero-extended planes are zeroed
// once per CTA and never written outside the image; clamp-extended planes get their outside copies from the inside
// quad next to the border, which stores its edge values outwards (up to R rows, one quad sideways, and the corners).
template <class G, int R>
__device__ __forceinline__ void replicate_out(const G& g, const Quad& q, const Plane<G, R>& dst, const float (&v)[4]) {
    const bool L = q.w == 0 && q.c >= 4, Rt = q.w + 4 == g.W && q.c + 8 <= G::P, U = q.h == 0, D = q.h == g.H - 1;
    const float lv[4] = {v[0], v[0], v[0], v[0]}, rv[4] = {v[3], v[3], v[3], v[3]};
    if (L) st4(dst.lrc(q.r, q.c - 4), lv);
    if (Rt) st4(dst.lrc(q.r, q.c + 4), rv);
#pragma unroll
    for (int k = 1; k <= R; ++k) {
        if (U && q.r - k >= 0) {
            st4(dst.lrc(q.r - k, q.c), v);
            if (L) st4(dst.lrc(q.r - k, q.c - 4), lv);
            if (Rt) st4(dst.lrc(q.r - k, q.c + 4), rv);
        }
        if (D && q.r + k < G::rows(R)) {
            st4(dst.lrc(q.r + k, q.c), v);
            if (L) st4(dst.lrc(q.r + k, q.c - 4), lv);
            if (Rt) st4(dst.lrc(q.r + k, q.c + 4), rv);
        }
    }
}

// ------------------------------------------------------------------ loads from global memory
// plane[h,w] = src[cl(h), cl(w)]  (CLAMP) or src[h,w] inside / 0 outside (zero-extended)
template <bool CLAMP, class G, int R>
__device__ __forceinline__ void load_plane(const G& g, const Plane<G, R>& dst, const float* __restrict__ src) {
    const bool vec = (g.W & 3) == 0;
    TILE_LOOP_NT(G::NT, i, G::items(R)) {
        const Quad q = quad_of<G, R>(g, i);
        float v[4];
        if (vec && q.fast) {
            ld4(src + (size_t)q.h * g.W + q.w, v);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CLAMP) v[j] = src[(size_t)glr_clampi(q.h, 0, g.H - 1) * g.W + glr_clampi(q.w + j, 0, g.W - 1)];
                else v[j] = g.inside(q.h, q.w + j) ? src[(size_t)q.h * g.W + q.w + j] : 0.f;
            }
        }
        st4(dst.lrc(q.r, q.c), v);
    }
}
template <class G, int R>
__device__ __forceinline__ void load_weights(const G& g, const WPl<G, R>& w, const float* __restrict__ wsrc) {
#pragma unroll
    for (int e = 0; e < 4; ++e) load_plane<false>(g, w.e[e], wsrc + (size_t)e * g.H * g.W);
}

// symmetric GTV coefficients: cR[q] = wR[q]^2 + wL[q+(0,1)]^2, cD[q] = wD[q]^2 + wU[q+(1,0)]^2 (0 outside Omega),
// computed straight from global memory into two planes of halo R.
template <class G, int R>
__device__ __forceinline__ void load_gtv_coeffs(const G& g, const Plane<G, R>& cR, const Plane<G, R>& cD,
                                                const float* __restrict__ wsrc) {
    const size_t HW = (size_t)g.H * g.W;
    const bool vec = (g.W & 3) == 0;
    TILE_LOOP_NT(G::NT, i, G::items(R)) {
        const Quad q = quad_of<G, R>(g, i);
        float vr[4] = {0.f, 0.f, 0.f, 0.f}, vd[4] = {0.f, 0.f, 0.f, 0.f};
        const size_t o = (size_t)q.h * g.W + q.w;
        if (vec && q.fast) {
            float wr[4], wd[4], wl[4], wu[4] = {0.f, 0.f, 0.f, 0.f};
            ld4(wsrc + 2 * HW + o, wr); ld4(wsrc + 3 * HW + o, wd); ld4(wsrc + 1 * HW + o, wl);
            if (q.h + 1 < g.H) ld4(wsrc + o + g.W, wu);
            const float wl4 = q.w + 4 < g.W ? wsrc[1 * HW + o + 4] : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wln = j == 3 ? wl4 : wl[j + 1];
                vr[j] = wr[j] * wr[j] + wln * wln;
                vd[j] = wd[j] * wd[j] + wu[j] * wu[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = q.w + j;
                if (!g.inside(q.h, w)) continue;
                const float wr = wsrc[2 * HW + o + j], wd = wsrc[3 * HW + o + j];
                const float wl = w + 1 < g.W ? wsrc[1 * HW + o + j + 1] : 0.f;
                const float wu = q.h + 1 < g.H ? wsrc[0 * HW + o + j + g.W] : 0.f;
                vr[j] = wr * wr + wl * wl;
                vd[j] = wd * wd + wu * wu;
            }
        }
        st4(cR.lrc(q.r, q.c), vr);
        st4(cD.lrc(q.r, q.c), vd);
    }
}

// Fused load of one stage input: fine plane (halo 3, clamp- or zero-extended) AND its 2x2 mean (coarse halo 3).
// One item = a pair of fine rows x one fine quad of the (+)6 region the coarse chain needs: two float4 global
// loads give two pooled values and (inside the (+)3 region) two rows of the fine plane.
// FN maps the loaded value(s) to the plane value: f(a, b) with a from src0 and b from src1 (b = 0 if src1 null).
template <bool CLAMP, class GF, class GC, class FN>
__device__ __forceinline__ void load_fine_and_pooled(const GF& gf, const GC& gc, const Plane<GF, 3>& fine,
                                                     const Plane<GC, 3>& coarse, const float* __restrict__ src0,
                                                     const float* __restrict__ src1, FN fn) {
    constexpr int NFQ = GF::NQ + 2;                  // fine quads of the (+)6 region: columns [w0-8, w0+TC+8)
    constexpr int NPAIR = GC::rows(3);               // fine row pairs == coarse rows
    const bool vec = (gf.W & 3) == 0;
    TILE_LOOP_NT(GF::NT, i, NPAIR * NFQ) {
        const int k = i / NFQ, fq = i % NFQ;
        const int h = gf.h0 - 6 + 2 * k, w = gf.w0 - 8 + 4 * fq;       // first fine pixel of the item
        const int hc = gc.h0 - 3 + k, wc = gc.w0 - 4 + 2 * fq;          // first coarse pixel
        float a[4], b[4], pooled[2];
        if (vec && h >= 0 && h + 1 < gf.H && w >= 0 && w + 3 < gf.W) {
            const size_t o = (size_t)h * gf.W + w;
            float t0[4], t1[4], u0[4] = {0.f, 0.f, 0.f, 0.f}, u1[4] = {0.f, 0.f, 0.f, 0.f};
            ld4(src0 + o, t0); ld4(src0 + o + gf.W, t1);
            if (src1) { ld4(src1 + o, u0); ld4(src1 + o + gf.W, u1); }
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = fn(t0[j], u0[j]); b[j] = fn(t1[j], u1[j]); }
            pooled[0] = 0.25f * (a[0] + a[1] + b[0] + b[1]);
            pooled[1] = 0.25f * (a[2] + a[3] + b[2] + b[3]);
        } else {
            auto val = [&](int hh, int ww) -> float {
                const size_t o = (size_t)hh * gf.W + ww;
                return fn(src0[o], src1 ? src1[o] : 0.f);
            };
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CLAMP) {
                    a[j] = val(glr_clampi(h, 0, gf.H - 1), glr_clampi(w + j, 0, gf.W - 1));
                    b[j] = val(glr_clampi(h + 1, 0, gf.H - 1), glr_clampi(w + j, 0, gf.W - 1));
                } else {
                    a[j] = gf.inside(h, w + j) ? val(h, w + j) : 0.f;
                    b[j] = gf.inside(h + 1, w + j) ? val(h + 1, w + j) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (CLAMP) {   // the coarse plane is clamp-extended in COARSE coordinates: mean of the block of cl(hc, wc)
                    const int y = 2 * glr_clampi(hc, 0, gc.H - 1), x = 2 * glr_clampi(wc + j, 0, gc.W - 1);
                    pooled[j] = 0.25f * (val(y, x) + val(y, x + 1) + val(y + 1, x) + val(y + 1, x + 1));
                } else {
                    pooled[j] = 0.25f * (a[2 * j] + a[2 * j + 1] + b[2 * j] + b[2 * j + 1]);   // zero extension commutes
                }
            }
        }
        *reinterpret_cast<float2*>(coarse.lrc(k, 2 * fq)) = make_float2(pooled[0], pooled[1]);
        if (fq >= 1 && fq <= GF::NQ) {
            const int r0 = 2 * k - 3, c = 4 * fq - 4;
            if (r0 >= 0 && r0 < GF::rows(3)) st4(fine.lrc(r0, c), a);
            if (r0 + 1 >= 0 && r0 + 1 < GF::rows(3)) st4(fine.lrc(r0 + 1, c), b);
        }
    }
}

// ------------------------------------------------------------------ asynchronous staging (cp.async)
// The next channel's input is copied global -> shared with cp.async (LDGSTS, no registers, no stall) into a RAW
// buffer while the current channel computes: rows [h0-6, h0+TR+6) x columns [w0-8, w0+TC+8), pitch TC+16, holding
// src[cl(h), cl(w)] (clamped addresses keep every copy in bounds; zero-extension is applied by the consumer).
template <class GF>
struct Raw {
    static constexpr int ROWS = GF::TR + 12, P = GF::TC + 16, NQ = P / 4, FLOATS = ROWS * P;
};
#ifdef GLRGTV_EMU
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    GLR_CHECK_ALIGN(dst, 16); GLR_CHECK_ALIGN(src, 16);
    if (emu_async_late) { emu_async_push(dst, src, 4); return; }
    for (int j = 0; j < 4; ++j) dst[j] = src[j];
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    if (emu_async_late) { emu_async_push(dst, src, 1); return; }
    *dst = *src;
}
__device__ __forceinline__ void cp_async_commit() { if (emu_async_late) emu_async_commit(); }
__device__ __forceinline__ void cp_async_wait_all() { if (emu_async_late) emu_async_wait(0); }
template <int N> __device__ __forceinline__ void cp_async_wait_pending() { if (emu_async_late) emu_async_wait(N); }
#else
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// wait until at most N of this thread's most recent commit groups are still in flight
template <int N> __device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// issue (do not wait for) the copy of one channel plane's (+)6 region into `raw`
template <class GF>
__device__ __forceinline__ void async_stage_raw(const GF& gf, float* raw, const float* __restrict__ src) {
    using R = Raw<GF>;
    const bool vec = (gf.W & 3) == 0;
    TILE_LOOP_NT(GF::NT, i, R::ROWS * R::NQ) {
        const int r = i / R::NQ, q = i % R::NQ;
        const int h = glr_clampi(gf.h0 - 6 + r, 0, gf.H - 1), w = gf.w0 - 8 + 4 * q;
        const float* row = src + (size_t)h * gf.W;
        float* dst = raw + r * R::P + 4 * q;
        if (vec && w >= 0 && w + 3 < gf.W) {
            cp_async16(dst, row + w);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async4(dst + j, row + glr_clampi(w + j, 0, gf.W - 1));
        }
    }
}

// Consume a staged raw buffer: fine plane (halo 3, clamp- or zero-extended) AND its 2x2 mean (coarse halo 3).
// One item = a pair of fine rows x one fine quad of the (+)6 region.  FN maps the staged value(s) to the plane
// value: f(a, b) with a from raw0 and b from raw1 (b = 0 if raw1 is null).
template <bool CLAMP, class GF, class GC, class FN>
__device__ __forceinline__ void consume_raw(const GF& gf, const GC& gc, const Plane<GF, 3>& fine, const Plane<GC, 3>& coarse,
                                            const float* raw0, const float* raw1, FN fn) {
    using R = Raw<GF>;
    constexpr int NPAIR = GC::rows(3);               // fine row pairs == coarse rows
    TILE_LOOP_NT(GF::NT, i, NPAIR * R::NQ) {
        const int k = i / R::NQ, fq = i % R::NQ;
        const int h = gf.h0 - 6 + 2 * k, w = gf.w0 - 8 + 4 * fq;       // first fine pixel of the item
        const int hc = gc.h0 - 3 + k, wc = gc.w0 - 4 + 2 * fq;          // first coarse pixel
        const int o = 2 * k * R::P + 4 * fq;
        float a[4], b[4], t0[4], t1[4], u0[4] = {0.f, 0.f, 0.f, 0.f}, u1[4] = {0.f, 0.f, 0.f, 0.f}, pooled[2];
        ld4(raw0 + o, t0); ld4(raw0 + o + R::P, t1);
        if (raw1) { ld4(raw1 + o, u0); ld4(raw1 + o + R::P, u1); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = fn(t0[j], u0[j]); b[j] = fn(t1[j], u1[j]); }
        const bool inside = h >= 0 && h + 1 < gf.H && w >= 0 && w + 3 < gf.W;
        if (!CLAMP && !inside) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!gf.inside(h, w + j)) a[j] = 0.f;
                if (!gf.inside(h + 1, w + j)) b[j] = 0.f;
            }
        }
        pooled[0] = 0.25f * (a[0] + a[1] + b[0] + b[1]);
        pooled[1] = 0.25f * (a[2] + a[3] + b[2] + b[3]);
        if (CLAMP && !inside) {
            // the coarse plane is clamp-extended in COARSE coordinates: mean of the 2x2 block of cl(hc, wc)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int y = 2 * glr_clampi(hc, 0, gc.H - 1) - (gf.h0 - 6), x = 2 * glr_clampi(wc + j, 0, gc.W - 1) - (gf.w0 - 8);
                const float* p0 = raw0 + y * R::P + x;
                pooled[j] = 0.25f * (fn(p0[0], 0.f) + fn(p0[1], 0.f) + fn(p0[R::P], 0.f) + fn(p0[R::P + 1], 0.f));
            }
        }
        *reinterpret_cast<float2*>(coarse.lrc(k, 2 * fq)) = make_float2(pooled[0], pooled[1]);
        if (fq >= 1 && fq <= GF::NQ) {
            const int r0 = 2 * k - 3, c = 4 * fq - 4;
            if (r0 >= 0 && r0 < GF::rows(3)) st4(fine.lrc(r0, c), a);
            if (r0 + 1 >= 0 && r0 + 1 < GF::rows(3)) st4(fine.lrc(r0 + 1, c), b);
        }
    }
}

// ------------------------------------------------------------------ forward stages (one quad each)
__device__ __forceinline__ float s_elem(const float* c, int P, const StatsTaps k) {
    return k.kc * c[0] + k.kr * c[1] + k.kd * c[P] + k.ku * c[-P] + k.kl * c[-1];
}
// S with two tap sets from one read: clamp-extended src (halo RS) -> clamp-extended dA, dB (halo RD)
template <bool TWO, class G, int RD, int RS>
__device__ __forceinline__ void q_S(const G& g, const Quad& q, const Plane<G, RD>& dA, const StatsTaps kA,
                                    const Plane<G, RD>& dB, const StatsTaps kB, const Plane<G, RS>& src) {
    static_assert(RS >= RD + 1, "halo");
    if (!G::GEN && !q.fast) return;
    float a[4], b[4];
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(src.lrc(q.r + RS - RD, q.c), n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] = kA.kc * n.c[j] + kA.kr * n.Rr(j) + kA.kd * n.d[j] + kA.ku * n.u[j] + kA.kl * n.L(j);
            if (TWO) b[j] = kB.kc * n.c[j] + kB.kr * n.Rr(j) + kB.kd * n.d[j] + kB.ku * n.u[j] + kB.kl * n.L(j);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* p = src.at(g, glr_clampi(q.h, 0, g.H - 1), glr_clampi(q.w + j, 0, g.W - 1));
            a[j] = s_elem(p, G::P, kA);
            if (TWO) b[j] = s_elem(p, G::P, kB);
        }
    }
    st4(dA.lrc(q.r, q.c), a);
    if (TWO) st4(dB.lrc(q.r, q.c), b);
    if (!G::GEN && q.border) {
        replicate_out(g, q, dA, a);
        if (TWO) replicate_out(g, q, dB, b);
    }
}

// L: s clamp-extended (halo RS) -> dst zero-extended (halo RD): s - sum_e w_e s[n_e]; weights halo RW >= RD
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void q_L(const G& g, const Quad& q, const Plane<G, RD>& dst, const Plane<G, RS>& s,
                                    const WPl<G, RW>& w) {
    if (!G::GEN && !q.fast) return;
    float v[4];
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(s.lrc(q.r + RS - RD, q.c), n);
        float w0[4], w1[4], w2[4], w3[4];
        const int rw = q.r + RW - RD;
        ld4(w.e[0].lrc(rw, q.c), w0); ld4(w.e[1].lrc(rw, q.c), w1); ld4(w.e[2].lrc(rw, q.c), w2); ld4(w.e[3].lrc(rw, q.c), w3);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = n.c[j] - (w0[j] * n.u[j] + w1[j] * n.L(j) + w2[j] * n.Rr(j) + w3[j] * n.d[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = 0.f;
            if (g.inside(q.h, q.w + j)) {
                const float* p = s.at(g, q.h, q.w + j);
                v[j] = p[0] - (*w.e[0].at(g, q.h, q.w + j) * p[-G::P] + *w.e[1].at(g, q.h, q.w + j) * p[-1] +
                               *w.e[2].at(g, q.h, q.w + j) * p[1] + *w.e[3].at(g, q.h, q.w + j) * p[G::P]);
            }
        }
    }
    st4(dst.lrc(q.r, q.c), v);
}

// phi(t) = 2*soft(t,G) - t = t - 2*clamp(t,-G,G)   (V1X0:765-777: epsilon - bias);  phi'(t) = |t|>G ? 1 : -1
__device__ __forceinline__ float glr_phi(float t, float G) { return t - 2.f * fminf(fmaxf(t, -G), G); }
__device__ __forceinline__ float glr_dphi(float t, float G) { return fabsf(t) > G ? 1.f : -1.f; }

// linear GTV core  o = Ct C s = sum_n c_n (s[q]-s[n])  (self-adjoint): s clamp-extended (halo RS),
// coefficient planes halo RC, dst zero-extended (halo RD).  (SURVEY B.5/B.6 combined.)
template <class G, int RD, int RS, int RC>
__device__ __forceinline__ void q_gtv_lin(const G& g, const Quad& q, const Plane<G, RD>& dst, const Plane<G, RS>& s,
                                          const Plane<G, RC>& cR, const Plane<G, RC>& cD) {
    static_assert(RC >= RD + 1 && RS >= RD + 1, "halo");
    if (!G::GEN && !q.fast) return;
    float v[4];
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(s.lrc(q.r + RS - RD, q.c), n);
        float cr[4], cd[4], cu[4];
        const float* pr = cR.lrc(q.r + RC - RD, q.c);
        ld4(pr, cr); ld4(cD.lrc(q.r + RC - RD, q.c), cd); ld4(cD.lrc(q.r + RC - RD - 1, q.c), cu);
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
            if (g.inside(q.h, q.w + j)) {
                const float* p = s.at(g, q.h, q.w + j);
                const float* a = cR.at(g, q.h, q.w + j);
                const float* b = cD.at(g, q.h, q.w + j);
                v[j] = a[0] * (p[0] - p[1]) + a[-1] * (p[0] - p[-1]) + b[0] * (p[0] - p[G::P]) + b[-G::P] * (p[0] - p[-G::P]);
            }
        }
    }
    st4(dst.lrc(q.r, q.c), v);
}

// raw weights of a quad and of the edges pointing at it: own[e][j] = w_e[q_j]; in[e][j] = weight of the edge from the
// neighbour in direction e back to q_j (edge D of the upper neighbour, R of the left, L of the right, U of the lower)
struct QW {
    float own[4][4], in[4][4];
};
template <class G, int RW>
__device__ __forceinline__ void ld_qw(const WPl<G, RW>& w, int rw, int c, QW& o) {
#pragma unroll
    for (int e = 0; e < 4; ++e) ld4(w.e[e].lrc(rw, c), o.own[e]);
    ld4(w.e[3].lrc(rw - 1, c), o.in[0]);
    ld4(w.e[0].lrc(rw + 1, c), o.in[3]);
    const float wr_m1 = *w.e[2].lrc(rw, c - 1), wl_p4 = *w.e[1].lrc(rw, c + 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.in[1][j] = j == 0 ? wr_m1 : o.own[2][j - 1];
        o.in[2][j] = j == 3 ? wl_p4 : o.own[1][j + 1];
    }
}
// GTV core with raw weights (zero-extended, halo RW >= RD+1), linear and/or thresholded from one read:
//   o[q] = sum_n [ wa phi(wa d) + wb phi(wb d) ],  d = s[q]-s[n], wa = w_{q->n}[q], wb = w_{n->q}[n]
template <bool LIN, bool THR, class G, int RD, int RS, int RW>
__device__ __forceinline__ void q_gtv_raw(const G& g, const Quad& q, const Plane<G, RD>& dLin, const Plane<G, RD>& dThr,
                                          const Plane<G, RS>& s, const WPl<G, RW>& w, float Gam) {
    static_assert(RW >= RD + 1 && RS >= RD + 1, "halo");
    if (!G::GEN && !q.fast) return;
    float vl[4] = {0.f, 0.f, 0.f, 0.f}, vt[4] = {0.f, 0.f, 0.f, 0.f};
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(s.lrc(q.r + RS - RD, q.c), n);
        QW qw;
        ld_qw(w, q.r + RW - RD, q.c, qw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d[4] = {n.c[j] - n.u[j], n.c[j] - n.L(j), n.c[j] - n.Rr(j), n.c[j] - n.d[j]};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float wa = qw.own[e][j], wb = qw.in[e][j];
                if (LIN) vl[j] += (wa * wa + wb * wb) * d[e];
                if (THR) vt[j] += wa * glr_phi(wa * d[e], Gam) + wb * glr_phi(wb * d[e], Gam);
            }
        }
    } else {
        constexpr int P = G::P;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!g.inside(q.h, q.w + j)) continue;
            const float* p = s.at(g, q.h, q.w + j);
            const float *w0 = w.e[0].at(g, q.h, q.w + j), *w1 = w.e[1].at(g, q.h, q.w + j), *w2 = w.e[2].at(g, q.h, q.w + j),
                        *w3 = w.e[3].at(g, q.h, q.w + j);
            const float d[4] = {p[0] - p[-P], p[0] - p[-1], p[0] - p[1], p[0] - p[P]};
            const float wa[4] = {w0[0], w1[0], w2[0], w3[0]}, wb[4] = {w3[-P], w2[-1], w1[1], w0[P]};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (LIN) vl[j] += (wa[e] * wa[e] + wb[e] * wb[e]) * d[e];
                if (THR) vt[j] += wa[e] * glr_phi(wa[e] * d[e], Gam) + wb[e] * glr_phi(wb[e] * d[e], Gam);
            }
        }
    }
    if (LIN) st4(dLin.lrc(q.r, q.c), vl);
    if (THR) st4(dThr.lrc(q.r, q.c), vt);
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

// VJP of St wrt its input = "S with zero padding", for up to two tap sets from one read of the zero-extended src:
//   dZ (zero-extended)  = sZ * sum_t kZ_t g[p + o_t]     (feeds the L adjoint)
//   dC (clamp-extended) = sC * sum_t kC_t g[p + o_t]     (feeds the GTV core)
template <bool HAS_Z, bool HAS_C, class G, int RD, int RS>
__device__ __forceinline__ void q_Szero(const G& g, const Quad& q, const Plane<G, RD>& dZ, const StatsTaps kZ, float sZ,
                                        const Plane<G, RD>& dC, const StatsTaps kC, float sC, const Plane<G, RS>& src) {
    static_assert(RS >= RD + 1, "halo");
    if (!G::GEN && !q.fast) return;
    float z[4], c[4];
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(src.lrc(q.r + RS - RD, q.c), n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (HAS_Z) z[j] = sZ * (kZ.kc * n.c[j] + kZ.kr * n.Rr(j) + kZ.kd * n.d[j] + kZ.ku * n.u[j] + kZ.kl * n.L(j));
            if (HAS_C) c[j] = sC * (kC.kc * n.c[j] + kC.kr * n.Rr(j) + kC.kd * n.d[j] + kC.ku * n.u[j] + kC.kl * n.L(j));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (HAS_Z) z[j] = g.inside(q.h, q.w + j) ? sZ * s_elem(src.at(g, q.h, q.w + j), G::P, kZ) : 0.f;
            if (HAS_C) c[j] = sC * s_elem(src.at(g, glr_clampi(q.h, 0, g.H - 1), glr_clampi(q.w + j, 0, g.W - 1)), G::P, kC);
        }
    }
    if (HAS_Z) st4(dZ.lrc(q.r, q.c), z);
    if (HAS_C) st4(dC.lrc(q.r, q.c), c);
    if (!G::GEN && HAS_C && q.border) replicate_out(g, q, dC, c);
}

// VJP of L wrt its input, zero-extended: gs[q] = gl[q] - sum_n w_{n->q}[n] gl[n] - sum_{e: q+d_e outside} w_e[q] gl[q]
// gl zero-extended (halo RS >= RD+1), w zero-extended (halo RW >= RD+1).
template <class G, int RD, int RS, int RW>
__device__ __forceinline__ void q_L_adj(const G& g, const Quad& q, const Plane<G, RD>& dst, const Plane<G, RS>& gl,
                                        const WPl<G, RW>& w) {
    static_assert(RS >= RD + 1 && RW >= RD + 1, "halo");
    if (!G::GEN && !q.fast) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int h = q.h, x = q.w;
    if (!G::GEN || q.fast) {
        N5 n;
        ld_n5<G::P>(gl.lrc(q.r + RS - RD, q.c), n);
        const int rw = q.r + RW - RD;
        float wu[4], wd[4], wl[4], wr[4];
        ld4(w.e[3].lrc(rw - 1, q.c), wu);   // edge D of the upper neighbour points at q
        ld4(w.e[0].lrc(rw + 1, q.c), wd);   // edge U of the lower neighbour
        const float* pR = w.e[2].lrc(rw, q.c);  // edge R of the left neighbour: columns c-1 .. c+2
        const float* pL = w.e[1].lrc(rw, q.c);  // edge L of the right neighbour: columns c+1 .. c+4
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
                if (h == 0) self += *w.e[0].lrc(rw, q.c + j);
                if (x + j == 0) self += wl[j];
                if (x + j == g.W - 1) self += wr[j];
                if (h == g.H - 1) self += *w.e[3].lrc(rw, q.c + j);
                v[j] -= self * n.c[j];
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xx = x + j;
            if (!g.inside(h, xx)) continue;
            const float* p = gl.at(g, h, xx);
            constexpr int P = G::P;
            float rr = p[0] - (w.e[3].at(g, h, xx)[-P] * p[-P] + w.e[2].at(g, h, xx)[-1] * p[-1] +
                               w.e[1].at(g, h, xx)[1] * p[1] + w.e[0].at(g, h, xx)[P] * p[P]);
            float self = 0.f;
            if (h == 0) self += *w.e[0].at(g, h, xx);
            if (xx == 0) self += *w.e[1].at(g, h, xx);
            if (xx == g.W - 1) self += *w.e[2].at(g, h, xx);
            if (h == g.H - 1) self += *w.e[3].at(g, h, xx);
            v[j] = rr - self * p[0];
        }
    }
    st4(dst.lrc(q.r, q.c), v);
}

// VJP of S wrt its input at one pixel, from a zero-extended gs plane
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

// VJP wrt s of the GTV core with raw weights: linear part on goA and/or thresholded part on goB, summed into dst
//   gs[q] = sum_n (goA[q]-goA[n]) (wa^2 + wb^2) + sum_n (goB[q]-goB[n]) (wa^2 phi'(wa d) + wb^2 phi'(wb d)),  d = s[q]-s[n]
// goA, goB, s clamp-extended, w zero-extended; all halos >= RD+1.
template <bool LIN, bool THR, class G, int RD, int RG, int RS, int RW>
__device__ __forceinline__ void q_gtv_raw_adj(const G& g, const Quad& q, const Plane<G, RD>& dst, const Plane<G, RG>& goA,
                                              const Plane<G, RG>& goB, const Plane<G, RS>& s, const WPl<G, RW>& w, float Gam) {
    if (!G::GEN && !q.fast) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (!G::GEN || q.fast) {
        N5 na, nb, ns;
        if (LIN) ld_n5<G::P>(goA.lrc(q.r + RG - RD, q.c), na);
        if (THR) { ld_n5<G::P>(goB.lrc(q.r + RG - RD, q.c), nb); ld_n5<G::P>(s.lrc(q.r + RS - RD, q.c), ns); }
        QW qw;
        ld_qw(w, q.r + RW - RD, q.c, qw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float wa = qw.own[e][j], wb = qw.in[e][j];
                if (LIN) {
                    const float D = e == 0 ? na.c[j] - na.u[j] : e == 1 ? na.c[j] - na.L(j) : e == 2 ? na.c[j] - na.Rr(j) : na.c[j] - na.d[j];
                    v[j] += D * (wa * wa + wb * wb);
                }
                if (THR) {
                    const float D = e == 0 ? nb.c[j] - nb.u[j] : e == 1 ? nb.c[j] - nb.L(j) : e == 2 ? nb.c[j] - nb.Rr(j) : nb.c[j] - nb.d[j];
                    const float d = e == 0 ? ns.c[j] - ns.u[j] : e == 1 ? ns.c[j] - ns.L(j) : e == 2 ? ns.c[j] - ns.Rr(j) : ns.c[j] - ns.d[j];
                    v[j] += D * (wa * wa * glr_dphi(wa * d, Gam) + wb * wb * glr_dphi(wb * d, Gam));
                }
            }
        }
    } else {
        constexpr int P = G::P;
        const int offs[4] = {-P, -1, 1, P};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!g.inside(q.h, q.w + j)) continue;
            const float* ws[4] = {w.e[0].at(g, q.h, q.w + j), w.e[1].at(g, q.h, q.w + j), w.e[2].at(g, q.h, q.w + j),
                                  w.e[3].at(g, q.h, q.w + j)};
            const float wa[4] = {ws[0][0], ws[1][0], ws[2][0], ws[3][0]}, wb[4] = {ws[3][-P], ws[2][-1], ws[1][1], ws[0][P]};
            const float* pa = LIN ? goA.at(g, q.h, q.w + j) : nullptr;
            const float* pb = THR ? goB.at(g, q.h, q.w + j) : nullptr;
            const float* ps = THR ? s.at(g, q.h, q.w + j) : nullptr;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (LIN) v[j] += (pa[0] - pa[offs[e]]) * (wa[e] * wa[e] + wb[e] * wb[e]);
                if (THR) {
                    const float d = ps[0] - ps[offs[e]];
                    v[j] += (pb[0] - pb[offs[e]]) * (wa[e] * wa[e] * glr_dphi(wa[e] * d, Gam) + wb[e] * wb[e] * glr_dphi(wb[e] * d, Gam));
                }
            }
        }
    }
    st4(dst.lrc(q.r, q.c), v);
}
