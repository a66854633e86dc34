// stream.cuh - register-streaming "row walkers" for the fused block kernels (3x3 cross window).
//
// A WALKER is a group of GL lanes that owns one channel plane at one resolution and walks it top to bottom, one image
// row per step.  Lane l owns the QUAD of columns [4l, 4l+4).  Every intermediate of the operator chain
//     z --S--> s --{L | GTV core}--> l/o --St--> A z
// lives in registers as a rolling window of rows (the newest row plus the two rows above it), so vertical taps are
// register reads and horizontal taps are ONE warp shuffle per side and row (the scalars left / right of a quad).
// Stage k lags stage k-1 by one row: the chain S -> core -> St finishes row t-3 when row t is loaded.
//
// A walker wider than a warp (GL = 64, two warps side by side) exchanges the scalars at the warp seam through a
// double-buffered shared-memory mailbox; because every horizontal tap reads the CENTRE row of its window - a row that
// was produced in the PREVIOUS step - one block barrier per step is enough.
//
// Border semantics (tile.cuh states them for the plane kernels; they are the same here):
//   * clamp-extended rows (inputs of S, L, the GTV core): the row above row 0 is row 0, the scalar left of column 0 is
//     column 0 - selected at USE time;
//   * zero-extended rows (inputs of St): rows outside the image are produced as zeros, scalars outside are zero.
#pragma once
#include "tile.cuh"

struct Row {
    float v[4];
};
__device__ __forceinline__ Row row_zero() { Row r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.f; return r; }
__device__ __forceinline__ Row row_sel(bool c, const Row& a, const Row& b) {
    Row r;
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = c ? a.v[j] : b.v[j];
    return r;
}
__device__ __forceinline__ Row row_ld(const float* p) { Row r; ld4(p, r.v); return r; }
__device__ __forceinline__ Row row_ld_if(bool ok, const float* p) { return ok ? row_ld(p) : row_zero(); }

// which lane of its walker a thread is, and where the walker's row ends
struct LaneCtx {
    int col0;        // first column of this lane's quad
    int width;       // shuffle segment width = min(GL, 32)
    bool active;     // col0 < W
    bool first;      // col0 == 0        : the left scalar follows the image-border rule
    bool last;       // col0 + 4 == W    : the right scalar follows the image-border rule
    bool seam_l;     // first lane of the second warp of a 64-lane walker: left scalar comes from the mailbox
    bool seam_r;     // last lane of the first warp: right scalar comes from the mailbox
    const float* mb_rd;   // mailbox written in the previous step  [plane][2]
    float* mb_wr;         // mailbox of this step
};

// scalars left / right of the quad `c`.  ZERO: zero-extended row, else clamp-extended.  XW: walker spans two warps.
template <bool ZERO, bool XW>
__device__ __forceinline__ void nb_lr(const Row& c, float& l, float& r, const LaneCtx& lc, int plane) {
    const float up = __shfl_up_sync(0xffffffffu, c.v[3], 1, lc.width);
    const float dn = __shfl_down_sync(0xffffffffu, c.v[0], 1, lc.width);
    l = lc.first ? (ZERO ? 0.f : c.v[0]) : up;
    r = lc.last ? (ZERO ? 0.f : c.v[3]) : dn;
    if (XW) {
        if (lc.seam_l) l = lc.mb_rd[2 * plane + 1];
        if (lc.seam_r) r = lc.mb_rd[2 * plane + 0];
    }
}
// publish the seam scalars of a row that the next step reads as a centre row
template <bool XW>
__device__ __forceinline__ void mb_post(const Row& c, const LaneCtx& lc, int plane) {
    if (XW) {
        if (lc.seam_l) lc.mb_wr[2 * plane + 0] = c.v[0];
        if (lc.seam_r) lc.mb_wr[2 * plane + 1] = c.v[3];
    }
}

__device__ __forceinline__ float rowL(const Row& c, float l, int j) { return j == 0 ? l : c.v[j - 1]; }
__device__ __forceinline__ float rowR(const Row& c, float r, int j) { return j == 3 ? r : c.v[j + 1]; }

// S (V1X0:177-195): k_c c + k_R right + k_D down + k_U up + k_L left
__device__ __forceinline__ Row w_S(const StatsTaps k, const Row& c, const Row& u, const Row& d, float l, float r) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = k.kc * c.v[j] + k.kr * rowR(c, r, j) + k.kd * d.v[j] + k.ku * u.v[j] + k.kl * rowL(c, l, j);
    return o;
}
// St (V1X0:197-215) and, with the same taps, the VJP of S on a zero-extended row: k_c c + k_R left + k_D up + k_U down + k_L right
__device__ __forceinline__ Row w_St(const StatsTaps k, const Row& c, const Row& u, const Row& d, float l, float r) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.v[j] = k.kc * c.v[j] + k.kr * rowL(c, l, j) + k.kd * u.v[j] + k.ku * d.v[j] + k.kl * rowR(c, r, j);
    return o;
}
// L (V1X0:218-228): c - (w_U up + w_L left + w_R right + w_D down)
__device__ __forceinline__ Row w_L(const Row& c, const Row& u, const Row& d, float l, float r, const Row (&w)[4]) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        o.v[j] = c.v[j] - (w[0].v[j] * u.v[j] + w[1].v[j] * rowL(c, l, j) + w[2].v[j] * rowR(c, r, j) + w[3].v[j] * d.v[j]);
    return o;
}
// linear GTV core Ct C s with the symmetric coefficients cR[q] = wR[q]^2 + wL[q+(0,1)]^2, cD[q] = wD[q]^2 + wU[q+(1,0)]^2:
//   o = cR (s - s_R) + cR[left] (s - s_L) + cD (s - s_D) + cD[up] (s - s_U)
__device__ __forceinline__ Row w_core_lin(const Row& c, const Row& u, const Row& d, float l, float r, const Row& cr,
                                          float cr_left, const Row& cd, const Row& cu) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float cl_ = j == 0 ? cr_left : cr.v[j - 1];
        o.v[j] = cr.v[j] * (c.v[j] - rowR(c, r, j)) + cl_ * (c.v[j] - rowL(c, l, j)) + cd.v[j] * (c.v[j] - d.v[j]) +
                 cu.v[j] * (c.v[j] - u.v[j]);
    }
    return o;
}

// raw GTV weights around one quad of one row: own[e] = w_e[q]; in[e] = weight of the edge that points from the
// neighbour in direction e back at q (zero outside the image)
struct RawW {
    Row own[4], in[4];
};
// wsrc: the weight set of this (batch, graph) [4][H][W]; row r must be inside the image
__device__ __forceinline__ void ld_raw_w(const float* __restrict__ wsrc, int H, int W, int r, int col0, RawW& o) {
    const size_t HW = (size_t)H * W;
    const float* p = wsrc + (size_t)r * W + col0;
#pragma unroll
    for (int e = 0; e < 4; ++e) o.own[e] = row_ld(p + e * HW);
    o.in[0] = row_ld_if(r > 0, p + 3 * HW - W);         // edge D of the upper neighbour
    o.in[3] = row_ld_if(r + 1 < H, p + W);              // edge U of the lower neighbour
    const float wr_m1 = col0 > 0 ? p[2 * HW - 1] : 0.f;       // edge R of the left neighbour
    const float wl_p4 = col0 + 4 < W ? p[1 * HW + 4] : 0.f;   // edge L of the right neighbour
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o.in[1].v[j] = j == 0 ? wr_m1 : o.own[2].v[j - 1];
        o.in[2].v[j] = j == 3 ? wl_p4 : o.own[1].v[j + 1];
    }
}
// thresholded GTV core (V1X0:757-781 folded): o = sum_n [ wa phi(wa d) + wb phi(wb d) ], d = s - s_n
__device__ __forceinline__ Row w_core_thr(const Row& c, const Row& u, const Row& d, float l, float r, const RawW& w, float Gam) {
    Row o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float dd[4] = {c.v[j] - u.v[j], c.v[j] - rowL(c, l, j), c.v[j] - rowR(c, r, j), c.v[j] - d.v[j]};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float wa = w.own[e].v[j], wb = w.in[e].v[j];
            acc += wa * glr_phi(wa * dd[e], Gam) + wb * glr_phi(wb * dd[e], Gam);
        }
        o.v[j] = acc;
    }
    return o;
}

// ------------------------------------------------------------------ TMA bulk copies + mbarrier (sm_90+ / sm_100a)
// One elected thread moves whole image rows global -> shared with cp.async.bulk (the TMA unit; SASS: UBLKCP) and the
// bytes are counted on an mbarrier that every consumer polls.  The emulation build copies synchronously.
#ifdef GLRGTV_EMU
typedef float* smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(float* p) { return p; }
__device__ __forceinline__ void mbar_init(smem_addr_t, unsigned) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(smem_addr_t, unsigned) {}
__device__ __forceinline__ void bulk_g2s(smem_addr_t dst, const float* src, unsigned bytes, smem_addr_t bar_) {
    GLR_CHECK_ALIGN(dst, 16); GLR_CHECK_ALIGN(src, 16); GLR_CHECK_ALIGN((uintptr_t)bytes, 16);       // cp.async.bulk: 16-byte addresses and size
    if (emu_async_late) { emu_bulk_push(dst, src, bytes, bar_); return; }
    memcpy(dst, src, bytes);
}
__device__ __forceinline__ void mbar_wait(smem_addr_t bar, unsigned) { if (emu_async_late) emu_bulk_wait(bar); }
__device__ __forceinline__ smem_addr_t smem_advance(smem_addr_t a, int floats) { return a + floats; }
__device__ __forceinline__ void cp_async16_s(smem_addr_t dst, const float* src) {
    GLR_CHECK_ALIGN(dst, 16); GLR_CHECK_ALIGN(src, 16);
    if (emu_async_late) { emu_async_push(dst, src, 4); return; }
    for (int j = 0; j < 4; ++j) dst[j] = src[j];
}
#else
typedef unsigned smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(float* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ smem_addr_t smem_advance(smem_addr_t a, int floats) { return a + 4u * (unsigned)floats; }
// 16-byte cp.async with a precomputed shared-window address
__device__ __forceinline__ void cp_async16_s(smem_addr_t dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void mbar_init(smem_addr_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
// make the barrier initialisation (and earlier generic-proxy writes to shared memory) visible to the async proxy
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(smem_addr_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(smem_addr_t dst, const float* src, unsigned bytes, smem_addr_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(smem_addr_t bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
#endif
