// tile.cuh - shared-memory tile views and stencil stages of the fused block kernels (3x3 cross window).
//
// A CTA owns one (batch, graph) pair and one TH x TW tile of the fine grid, and walks the graph's F
// signal channels one after another.  Every intermediate of the operator chain lives in a shared-memory
// "view": a rectangle of GLOBAL pixel coordinates [oh, oh+nh) x [ow, ow+nw) (origin may be negative /
// beyond the image), stored row-major.  Two conventions for pixels outside the image Omega:
//   * clamp-extended views hold X[cl(p)]  (what a replicate-padded gather reads: S, L, C, P inputs);
//   * zero-extended  views hold 0         (what the transposed stencils St and Ct read).
// A clamp-extended stage output is produced by evaluating the stencil at the CLAMPED centre cl(p) on a
// clamp-extended input, which is exactly "value at the replicated position".
//
// Edge order of the cross window (V1X0:26-30, 42-49): e=0 U(-1,0), e=1 L(0,-1), e=2 R(0,1), e=3 D(1,0).
#pragma once
#include "common.cuh"

struct View {
    float* p;
    int oh, ow, nh, nw;
    __device__ __forceinline__ float& at(int h, int w) const { return p[(h - oh) * nw + (w - ow)]; }
    __device__ __forceinline__ int size() const { return nh * nw; }
};

__device__ __forceinline__ View make_view(float*& cursor, int oh, int ow, int nh, int nw) {
    View v{cursor, oh, ow, nh, nw};
    cursor += (nh * nw + 3) & ~3;  // keep 16-byte alignment of every view
    return v;
}

#define TILE_LOOP(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)

// dst[h,w] = plane[cl(h), cl(w)]   (clamp-extended load)
__device__ __forceinline__ void tile_load_clamped(const View& dst, const float* __restrict__ plane, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, w = dst.ow + i % dst.nw;
        dst.p[i] = plane[glr_clampi(h, 0, H - 1) * W + glr_clampi(w, 0, W - 1)];
    }
}
// dst[h,w] = plane[h,w] inside the image, 0 outside   (zero-extended load)
__device__ __forceinline__ void tile_load_zero(const View& dst, const float* __restrict__ plane, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, w = dst.ow + i % dst.nw;
        dst.p[i] = glr_inside(h, w, H, W) ? plane[h * W + w] : 0.f;
    }
}

// S: clamp-extended in -> clamp-extended out.  src must cover dst (+) 1.
__device__ __forceinline__ void tile_S(const View& dst, const View& src, const StatsTaps k, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = glr_clampi(dst.oh + i / dst.nw, 0, H - 1), w = glr_clampi(dst.ow + i % dst.nw, 0, W - 1);
        const float* c = &src.at(h, w);
        dst.p[i] = k.kc * c[0] + k.kr * c[1] + k.kd * c[src.nw] + k.ku * c[-src.nw] + k.kl * c[-1];
    }
}
// two S with different taps from one read of the input
__device__ __forceinline__ void tile_S2(const View& dA, const StatsTaps kA, const View& dB, const StatsTaps kB,
                                        const View& src, int H, int W) {
    TILE_LOOP(i, dA.size()) {
        int h = glr_clampi(dA.oh + i / dA.nw, 0, H - 1), w = glr_clampi(dA.ow + i % dA.nw, 0, W - 1);
        const float* c = &src.at(h, w);
        float vc = c[0], vr = c[1], vd = c[src.nw], vu = c[-src.nw], vl = c[-1];
        dA.p[i] = kA.kc * vc + kA.kr * vr + kA.kd * vd + kA.ku * vu + kA.kl * vl;
        dB.p[i] = kB.kc * vc + kB.kr * vr + kB.kd * vd + kB.ku * vu + kB.kl * vl;
    }
}
// 2x2 mean: fine clamp-extended src -> coarse clamp-extended dst (coarse coordinates, coarse image Hc x Wc)
__device__ __forceinline__ void tile_pool(const View& dst, const View& src, int Hc, int Wc) {
    TILE_LOOP(i, dst.size()) {
        int h = glr_clampi(dst.oh + i / dst.nw, 0, Hc - 1), w = glr_clampi(dst.ow + i % dst.nw, 0, Wc - 1);
        const float* c = &src.at(2 * h, 2 * w);
        dst.p[i] = 0.25f * (c[0] + c[1] + c[src.nw] + c[src.nw + 1]);
    }
}

// the four weight planes of one graph, one view each (same rectangle)
struct WViews {
    View e[4];
};
// weights are loaded ZERO-extended: an edge that leaves the image has weight 0 on the far side
__device__ __forceinline__ void tile_load_weights(const WViews& wv, const float* __restrict__ wplane, int H, int W) {
    const int HW = H * W;
#pragma unroll
    for (int e = 0; e < 4; ++e) tile_load_zero(wv.e[e], wplane + (size_t)e * HW, H, W);
}
__device__ __forceinline__ WViews make_wviews(float*& cursor, int oh, int ow, int nh, int nw) {
    WViews wv;
#pragma unroll
    for (int e = 0; e < 4; ++e) wv.e[e] = make_view(cursor, oh, ow, nh, nw);
    return wv;
}

// L: dst = s - sum_e w_e s[n_e] inside the image, 0 outside (zero-extended out; feeds St).
// s clamp-extended covering dst (+) 1; w covering dst.
__device__ __forceinline__ void tile_L(const View& dst, const View& s, const WViews& w, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, x = dst.ow + i % dst.nw;
        float r = 0.f;
        if (glr_inside(h, x, H, W)) {
            const float* c = &s.at(h, x);
            r = c[0] - (w.e[0].at(h, x) * c[-s.nw] + w.e[1].at(h, x) * c[-1] + w.e[2].at(h, x) * c[1] +
                        w.e[3].at(h, x) * c[s.nw]);
        }
        dst.p[i] = r;
    }
}

// phi(t) = 2*soft(t,G) - t = t - 2*clamp(t,-G,G)        (V1X0:765-777: epsilon - bias)
__device__ __forceinline__ float glr_phi(float t, float G) { return t - 2.f * fminf(fmaxf(t, -G), G); }

// GTV core  o = Ct phi(C s)  for the cross window, zero-extended out:
//   o[q] = sum_{n in N4(q)} [ wa*phi(wa*d) + wb*phi(wb*d) ],  d = s[q]-s[n], wa = w_{q->n}[q], wb = w_{n->q}[n]
// (SURVEY B.5/B.6 combined; the zero-extended weights and clamp-extended s make every border term vanish).
// THR=false is the linear case phi(t)=t.  s covers dst (+) 1, w covers dst (+) 1.
template <bool THR>
__device__ __forceinline__ float gtv_pair(float wa, float wb, float d, float G) {
    if (THR) return wa * glr_phi(wa * d, G) + wb * glr_phi(wb * d, G);
    return (wa * wa + wb * wb) * d;
}
template <bool THR>
__device__ __forceinline__ void tile_gtv_core(const View& dst, const View& s, const WViews& w, float G, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, x = dst.ow + i % dst.nw;
        float r = 0.f;
        if (glr_inside(h, x, H, W)) {
            const float* c = &s.at(h, x);
            float v = c[0];
            // neighbour U: q->n is edge 0 (U) at q, n->q is edge 3 (D) at n
            r = gtv_pair<THR>(w.e[0].at(h, x), w.e[3].at(h - 1, x), v - c[-s.nw], G);
            r += gtv_pair<THR>(w.e[1].at(h, x), w.e[2].at(h, x - 1), v - c[-1], G);
            r += gtv_pair<THR>(w.e[2].at(h, x), w.e[1].at(h, x + 1), v - c[1], G);
            r += gtv_pair<THR>(w.e[3].at(h, x), w.e[0].at(h + 1, x), v - c[s.nw], G);
        }
        dst.p[i] = r;
    }
}

// St at one pixel from a zero-extended view: sum_t k_t y[q - o_t]
__device__ __forceinline__ float tile_St_at(const View& y, const StatsTaps k, int h, int w) {
    const float* c = &y.at(h, w);
    return k.kc * c[0] + k.kr * c[-1] + k.kd * c[-y.nw] + k.ku * c[y.nw] + k.kl * c[1];
}

// ====================================================================================================
// adjoint (VJP) stages - SURVEY Appendix B.2-B.6.  For a unit offset d the adjoint of the clamped gather
// x[cl(p+d)] is   sum_{p: cl(p+d)=q} v[p] = v_zero[q-d] + [q+d outside] v[q]   (the border pixel also
// collects the tap that was replicated onto it).
// ====================================================================================================

// VJP of St wrt its input = "S with zero padding": out[p] = scale * sum_t k_t g[p + o_t], g zero-extended.
// CLAMP_OUT: store the value of the clamped centre (clamp-extended result, feeds the GTV core),
// otherwise 0 outside the image (zero-extended result, feeds the L adjoint).  g covers dst (+) 1.
template <bool CLAMP_OUT>
__device__ __forceinline__ void tile_Szero(const View& dst, const View& g, const StatsTaps k, float scale, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, w = dst.ow + i % dst.nw;
        if (CLAMP_OUT) {
            h = glr_clampi(h, 0, H - 1);
            w = glr_clampi(w, 0, W - 1);
        } else if (!glr_inside(h, w, H, W)) {
            dst.p[i] = 0.f;
            continue;
        }
        const float* c = &g.at(h, w);
        dst.p[i] = scale * (k.kc * c[0] + k.kr * c[1] + k.kd * c[g.nw] + k.ku * c[-g.nw] + k.kl * c[-1]);
    }
}

// VJP of L wrt its input, zero-extended out: gs[q] = gl[q] - sum_n w_{n->q}[n] gl[n] - sum_{e: q+d_e outside} w_e[q] gl[q]
// gl zero-extended covering dst (+) 1, w zero-extended covering dst (+) 1.
__device__ __forceinline__ void tile_L_adj(const View& dst, const View& gl, const WViews& w, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, x = dst.ow + i % dst.nw;
        float r = 0.f;
        if (glr_inside(h, x, H, W)) {
            const float* c = &gl.at(h, x);
            float v = c[0];
            r = v - (w.e[3].at(h - 1, x) * c[-gl.nw] + w.e[2].at(h, x - 1) * c[-1] + w.e[1].at(h, x + 1) * c[1] +
                     w.e[0].at(h + 1, x) * c[gl.nw]);
            float self = 0.f;
            if (h == 0) self += w.e[0].at(h, x);
            if (x == 0) self += w.e[1].at(h, x);
            if (x == W - 1) self += w.e[2].at(h, x);
            if (h == H - 1) self += w.e[3].at(h, x);
            r -= self * v;
        }
        dst.p[i] = r;
    }
}

// VJP of S wrt its input at one pixel, from a zero-extended gs
__device__ __forceinline__ float tile_S_adj_at(const View& gs, const StatsTaps k, int h, int w, int H, int W) {
    const float* c = &gs.at(h, w);
    float v = c[0];
    float r = k.kc * v + k.kr * c[-1] + k.kd * c[-gs.nw] + k.ku * c[gs.nw] + k.kl * c[1];
    float self = 0.f;
    if (w == W - 1) self += k.kr;
    if (h == H - 1) self += k.kd;
    if (h == 0) self += k.ku;
    if (w == 0) self += k.kl;
    return r + self * v;
}

// phi'(t): +1 where |t| > G, else -1   (the reference's strict comparisons, V1X0:691-703)
__device__ __forceinline__ float glr_dphi(float t, float G) { return fabsf(t) > G ? 1.f : -1.f; }

// VJP of the thresholded GTV core wrt s (added into dst, which already holds the linear part or 0):
//   gs[q] (+)= sum_n (go[q]-go[n]) (wa^2 phi'(wa d) + wb^2 phi'(wb d)),  d = s[q]-s[n]
// go, s clamp-extended covering dst (+) 1; w zero-extended covering dst (+) 1.
template <bool ACCUM>
__device__ __forceinline__ void tile_gtv_core_thr_adj(const View& dst, const View& go, const View& s, const WViews& w,
                                                      float G, int H, int W) {
    TILE_LOOP(i, dst.size()) {
        int h = dst.oh + i / dst.nw, x = dst.ow + i % dst.nw;
        float r = 0.f;
        if (glr_inside(h, x, H, W)) {
            const float* cg = &go.at(h, x);
            const float* cs = &s.at(h, x);
            const float gv = cg[0], sv = cs[0];
            float wa, wb, d;
            wa = w.e[0].at(h, x); wb = w.e[3].at(h - 1, x); d = sv - cs[-s.nw];
            r += (gv - cg[-go.nw]) * (wa * wa * glr_dphi(wa * d, G) + wb * wb * glr_dphi(wb * d, G));
            wa = w.e[1].at(h, x); wb = w.e[2].at(h, x - 1); d = sv - cs[-1];
            r += (gv - cg[-1]) * (wa * wa * glr_dphi(wa * d, G) + wb * wb * glr_dphi(wb * d, G));
            wa = w.e[2].at(h, x); wb = w.e[1].at(h, x + 1); d = sv - cs[1];
            r += (gv - cg[1]) * (wa * wa * glr_dphi(wa * d, G) + wb * wb * glr_dphi(wb * d, G));
            wa = w.e[3].at(h, x); wb = w.e[0].at(h + 1, x); d = sv - cs[s.nw];
            r += (gv - cg[go.nw]) * (wa * wa * glr_dphi(wa * d, G) + wb * wb * glr_dphi(wb * d, G));
            if (ACCUM) r += dst.p[i];
        } 
        dst.p[i] = r;
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
