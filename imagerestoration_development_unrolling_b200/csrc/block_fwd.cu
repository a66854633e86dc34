// block_fwd.cu - fused forward of LocalLowpassFilteringBlock / MixtureGTVGLR (V1X0:707-811, 985-988).
//
// Staged fusion: the unrolled solver is cut where a global dependency forces it, and each
// stage is ONE kernel that reads its operands once from HBM, keeps the whole operator chain of
// A(.) = I + e^mu0 St L S + e^ro0 St Ct C S + P^T[ e^mu1 ... + e^ro1 ... ]P  (V1X0:642-682)
// in shared memory, and writes only what later stages or the backward pass need:
//
//   k_block_weights x2   feat0 / feat1           -> wT0,wL0 / wT1,wL1        (V1X0:716-733)
//   MODE_BA              y                        -> bA = y + R_lin(y)         (V1X0:738-749)
//   MODE_X1              bA                       -> x1 = bA + a0 (bA - A bA)  (V1X0:751-753)
//   MODE_X2              x1, y                    -> bB, r1, x2                (V1X0:757-786)
//   MODE_X3              x2, bB, r1, x            -> out                       (V1X0:788-790, 985-988)
//
// with R(z) = e^ro0 St0 Ct0 phi(C0 S0 z) + P^T e^ro1 St1 Ct1 phi(C1 S1 P z), phi = identity (R_lin) or
// 2*soft-threshold - id (bB).  Tile = TH x TW fine pixels, input halo 6 (3 for the fine chain, 3 coarse
// pixels = 6 fine for the half-resolution chain).  Stage arithmetic lives in tile.cuh.
#include "tile.cuh"

enum { MODE_BA = 0, MODE_X1 = 1, MODE_X2 = 2, MODE_X3 = 3 };

struct BlockFwdArgs {
    glrgtv_shape s;
    glrgtv_block_params p;
    const float* z;      // stencil input: y | bA | x1 | x2
    const float* y;      // X2: y ; X3: x (skip path)
    const float* bB_in;  // X3
    const float* r1_in;  // X3
    const float *wT0, *wL0, *wT1, *wL1;
    float* out0;  // BA: bA | X1: x1 | X2: x2 | X3: out
    float* out1;  // X2: bB
    float* out2;  // X2: r1
};

// shared-memory layout (floats) of one forward stage; 4 guard floats on either side (quad 0 / NQ-1 read the scalar
// left / right of their row, which for the first / last plane row lies just outside the plane)
template <int MODE, int TH, int TW, int NT, bool GEN>
struct FwdLayout {
    using GF = Geo<TH, TW, NT, GEN>;
    using GC = Geo<TH / 2, TW / 2, NT, GEN>;
    static constexpr bool GLR = MODE != MODE_BA, THR = MODE == MODE_X2;
    static constexpr int zf = 4;
    static constexpr int sA = zf + GF::floats(3);
    static constexpr int sB = sA + (GLR ? GF::floats(2) : 0);
    static constexpr int lA = sB + GF::floats(2);
    static constexpr int oB = lA + (GLR ? GF::floats(1) : 0);
    static constexpr int oT = oB + GF::floats(1);
    static constexpr int pz = oT + (THR ? GF::floats(1) : 0);
    static constexpr int sA1 = pz + GC::floats(3);
    static constexpr int sB1 = sA1 + (GLR ? GC::floats(2) : 0);
    static constexpr int lA1 = sB1 + GC::floats(2);
    static constexpr int oB1 = lA1 + (GLR ? GC::floats(1) : 0);
    static constexpr int oT1 = oB1 + GC::floats(1);
    // weights: GLR raw (halo 1); GTV as symmetric coefficient planes, or raw (halo 2) where the threshold needs them
    static constexpr int wL0 = oT1 + (THR ? GC::floats(1) : 0);
    static constexpr int cR0 = wL0 + (GLR ? 4 * GF::floats(1) : 0);
    static constexpr int cD0 = cR0 + (THR ? 0 : GF::floats(2));
    static constexpr int wT0 = cD0 + (THR ? 0 : GF::floats(2));
    static constexpr int wL1 = wT0 + (THR ? 4 * GF::floats(2) : 0);
    static constexpr int cR1 = wL1 + (GLR ? 4 * GC::floats(1) : 0);
    static constexpr int cD1 = cR1 + (THR ? 0 : GC::floats(2));
    static constexpr int wT1 = cD1 + (THR ? 0 : GC::floats(2));
    static constexpr int raw = wT1 + (THR ? 4 * GC::floats(2) : 0) + 4;   // cp.async staging of the next channel
    static constexpr int total = raw + Raw<GF>::FLOATS;
};

// resident CTAs per SM the shared-memory footprint allows (the register budget follows it)
#ifdef GLR_FWD_MINB
constexpr int fwd_min_blocks(int) { return GLR_FWD_MINB; }
#else
constexpr int fwd_min_blocks(int mode) { return mode == MODE_BA ? 4 : 2; }
#endif

template <int MODE, int TH, int TW, int NT, bool GEN>
__global__ void __launch_bounds__(NT, fwd_min_blocks(MODE)) k_block_stage(BlockFwdArgs a) {
    GLR_SMEM_DECL(smem);
    using LY = FwdLayout<MODE, TH, TW, NT, GEN>;
    using GF = typename LY::GF;
    using GC = typename LY::GC;
    constexpr bool GLR = LY::GLR, THR = LY::THR;
    const int H = a.s.H, W = a.s.W, F = a.s.F, G = a.s.G;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, b = plane / G;
    GF gf; gf.H = H; gf.W = W; gf.h0 = (tile / tiles_w) * TH; gf.w0 = (tile % tiles_w) * TW;
    GC gc; gc.H = H / 2; gc.W = W / 2; gc.h0 = gf.h0 / 2; gc.w0 = gf.w0 / 2;
    const size_t HW = (size_t)H * W, HWc = HW / 4;

    auto zf = plane_at<GF, 3>(smem, LY::zf);
    auto sA = plane_at<GF, 2>(smem, LY::sA);
    auto sB = plane_at<GF, 2>(smem, LY::sB);
    auto lA = plane_at<GF, 1>(smem, LY::lA);
    auto oB = plane_at<GF, 1>(smem, LY::oB);
    auto oT = plane_at<GF, 1>(smem, LY::oT);
    auto pz = plane_at<GC, 3>(smem, LY::pz);
    auto sA1 = plane_at<GC, 2>(smem, LY::sA1);
    auto sB1 = plane_at<GC, 2>(smem, LY::sB1);
    auto lA1 = plane_at<GC, 1>(smem, LY::lA1);
    auto oB1 = plane_at<GC, 1>(smem, LY::oB1);
    auto oT1 = plane_at<GC, 1>(smem, LY::oT1);
    auto wL0 = wplanes_at<GF, 1>(smem, LY::wL0);
    auto cR0 = plane_at<GF, 2>(smem, LY::cR0);
    auto cD0 = plane_at<GF, 2>(smem, LY::cD0);
    auto wT0 = wplanes_at<GF, 2>(smem, LY::wT0);
    auto wL1 = wplanes_at<GC, 1>(smem, LY::wL1);
    auto cR1 = plane_at<GC, 2>(smem, LY::cR1);
    auto cD1 = plane_at<GC, 2>(smem, LY::cD1);
    auto wT1 = wplanes_at<GC, 2>(smem, LY::wT1);

    // ---- per-graph scalars
    const float aT0 = expf(a.p.ro0[g]), aT1 = expf(a.p.ro1[g]);
    const float aL0 = GLR ? expf(a.p.mu0[g]) : 0.f, aL1 = GLR ? expf(a.p.mu1[g]) : 0.f;
    const float G0 = THR ? expf(a.p.gamma0[g]) : 0.f, G1 = THR ? expf(a.p.gamma1[g]) : 0.f;
    float alpha = 0.f, beta2 = 0.f, s0 = 0.f, s1 = 1.f;
    if (MODE == MODE_X1) alpha = a.p.alpha[0 * G + g];
    if (MODE == MODE_X2) alpha = a.p.alpha[1 * G + g];
    if (MODE == MODE_X3) {
        alpha = a.p.alpha[2 * G + g];
        beta2 = a.p.beta[2 * G + g];
        if (a.p.skip) { s0 = a.p.skip[0]; s1 = a.p.skip[1]; }
    }
    const bool has_skip = MODE == MODE_X3 && a.p.skip != nullptr;

    // branch-free borders: the zero-extended planes are zeroed once, outside quads are never written afterwards
    if (!GEN) {
        const float z4[4] = {0.f, 0.f, 0.f, 0.f};
        TILE_LOOP_NT(NT, i, (LY::wL0 - LY::zf) / 4) st4(smem + LY::zf + 4 * i, z4);
    }

    // ---- weights of this graph (shared by its F channels)
    const size_t wplane = (size_t)plane * 4;
    if (THR) {
        load_weights(gf, wT0, a.wT0 + wplane * HW);
        load_weights(gc, wT1, a.wT1 + wplane * HWc);
    } else {
        load_gtv_coeffs(gf, cR0, cD0, a.wT0 + wplane * HW);
        load_gtv_coeffs(gc, cR1, cD1, a.wT1 + wplane * HWc);
    }
    if (GLR) {
        load_weights(gf, wL0, a.wL0 + wplane * HW);
        load_weights(gc, wL1, a.wL1 + wplane * HWc);
    }
    const bool vec = (W & 3) == 0;

    float* raw = smem + LY::raw;
    const size_t plane0 = ((size_t)b * G * F + (size_t)g * F) * HW;
    async_stage_raw(gf, raw, a.z + plane0);      // channel 0; later channels are staged while their predecessor computes
    cp_async_commit();

    for (int f = 0; f < F; ++f) {
        const int c = g * F + f;
        const size_t off = plane0 + (size_t)f * HW;
        const StatsTaps kT0 = glr_load_taps(a.p.gtv0.stats, c), kT1 = glr_load_taps(a.p.gtv1.stats, c);
        StatsTaps kL0 = kT0, kL1 = kT1;
        if (GLR) { kL0 = glr_load_taps(a.p.glr0.stats, c); kL1 = glr_load_taps(a.p.glr1.stats, c); }

        cp_async_wait_all();
        __syncthreads();  // this channel's raw input has landed; the previous channel's epilogue is done with the planes
        // phase 0: stage input on the tile (+) 3 and its 2x2 mean on the coarse tile (+) 3, from the staged raw buffer
        consume_raw<true>(gf, gc, zf, pz, raw, (const float*)nullptr, [](float v, float) { return v; });
        __syncthreads();
        if (f + 1 < F) {
            async_stage_raw(gf, raw, a.z + off + HW);
            cp_async_commit();
        }
        // phase 1: S at both resolutions
        TILE_LOOP_NT(NT, i, GF::items(2)) {
            const Quad q = quad_of<GF, 2>(gf, i);
            q_S<GLR>(gf, q, GLR ? sA : sB, GLR ? kL0 : kT0, sB, kT0, zf);
        }
        TILE_LOOP_REV(NT, i, GC::items(2)) {
            const Quad q = quad_of<GC, 2>(gc, i);
            q_S<GLR>(gc, q, GLR ? sA1 : sB1, GLR ? kL1 : kT1, sB1, kT1, pz);
        }
        __syncthreads();
        // phase 2: L and the GTV cores at both resolutions
        TILE_LOOP_NT(NT, i, GF::items(1)) {
            const Quad q = quad_of<GF, 1>(gf, i);
            if (GLR) q_L(gf, q, lA, sA, wL0);
            if (THR) q_gtv_raw<true, true>(gf, q, oB, oT, sB, wT0, G0);
            else q_gtv_lin(gf, q, oB, sB, cR0, cD0);
        }
        TILE_LOOP_REV(NT, i, GC::items(1)) {
            const Quad q = quad_of<GC, 1>(gc, i);
            if (GLR) q_L(gc, q, lA1, sA1, wL1);
            if (THR) q_gtv_raw<true, true>(gc, q, oB1, oT1, sB1, wT1, G1);
            else q_gtv_lin(gc, q, oB1, sB1, cR1, cD1);
        }
        __syncthreads();
        // phase 3: fine St (+ the two coarse St values under each quad) and the stage epilogue
        TILE_LOOP_NT(NT, i, GF::TR * (GF::NQ - 2)) {
            const Quad q = tile_quad_of(gf, i);
            const int r = q.r, cq = q.c, h = q.h, w = q.w;
            if (h >= H || w >= W) continue;
            float zq[4], Az[4], rT[4];
            ld4(zf.lrc(r + 3, cq), zq);
            St_quad<GF::P>(oB.lrc(r + 1, cq), kT0, Az);
            const int rc = (r >> 1) + 1, cc = (cq >> 1) + 2;   // coarse row (halo-1 planes) / column under this quad
            float tc0 = aT1 * St_elem(oB1.lrc(rc, cc), GC::P, kT1), tc1 = aT1 * St_elem(oB1.lrc(rc, cc + 1), GC::P, kT1);
            if (GLR) {
                float gl_[4];
                St_quad<GF::P>(lA.lrc(r + 1, cq), kL0, gl_);
                tc0 += aL1 * St_elem(lA1.lrc(rc, cc), GC::P, kL1);
                tc1 += aL1 * St_elem(lA1.lrc(rc, cc + 1), GC::P, kL1);
#pragma unroll
                for (int j = 0; j < 4; ++j) Az[j] = aT0 * Az[j] + aL0 * gl_[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) Az[j] = aT0 * Az[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) Az[j] += zq[j] + 0.25f * (j < 2 ? tc0 : tc1);
            if (THR) {
                St_quad<GF::P>(oT.lrc(r + 1, cq), kT0, rT);
                const float t0 = aT1 * St_elem(oT1.lrc(rc, cc), GC::P, kT1), t1 = aT1 * St_elem(oT1.lrc(rc, cc + 1), GC::P, kT1);
#pragma unroll
                for (int j = 0; j < 4; ++j) rT[j] = aT0 * rT[j] + 0.25f * (j < 2 ? t0 : t1);
            }
            const size_t gi = off + (size_t)h * W + w;
            const bool full = !GEN || (vec && w + 3 < W);
            float o0[4], o1[4], o2[4], in0[4], in1[4], in2[4];
            // pointwise operands
            if (MODE == MODE_X2 || MODE == MODE_X3) {
                if (full) {
                    if (MODE == MODE_X2 || has_skip) ld4(a.y + gi, in0);
                    if (MODE == MODE_X3) { ld4(a.bB_in + gi, in1); ld4(a.r1_in + gi, in2); }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = w + j < W;
                        in0[j] = ok && (MODE == MODE_X2 || has_skip) ? a.y[gi + j] : 0.f;
                        in1[j] = ok && MODE == MODE_X3 ? a.bB_in[gi + j] : 0.f;
                        in2[j] = ok && MODE == MODE_X3 ? a.r1_in[gi + j] : 0.f;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == MODE_BA) {
                    o0[j] = Az[j];  // y + R_lin(y)
                } else if (MODE == MODE_X1) {
                    o0[j] = zq[j] + alpha * (zq[j] - Az[j]);
                } else if (MODE == MODE_X2) {
                    const float bB = in0[j] + rT[j], r1 = bB - Az[j];
                    o1[j] = bB;
                    o2[j] = r1;
                    o0[j] = zq[j] + alpha * r1;
                } else {
                    const float u2 = (in1[j] - Az[j]) + beta2 * in2[j];
                    const float x3 = zq[j] + alpha * u2;
                    o0[j] = has_skip ? s0 * in0[j] + s1 * x3 : x3;
                }
            }
            if (full) {
                st4(a.out0 + gi, o0);
                if (MODE == MODE_X2) { st4(a.out1 + gi, o1); st4(a.out2 + gi, o2); }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w + j < W) {
                        a.out0[gi + j] = o0[j];
                        if (MODE == MODE_X2) { a.out1[gi + j] = o1[j]; a.out2[gi + j] = o2[j]; }
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// edge weights of both operator families of one scale in one launch (V1X0:716-733).
// feat [B, 2C, H, W]: channels [0,C) -> GTV weights, [C,2C) -> GLR weights.  plane = (b, set, g).
// ---------------------------------------------------------------------------------------------------
template <int TH, int TW>
__global__ void __launch_bounds__(GLR_THREADS) k_block_weights(glrgtv_shape s, const float* __restrict__ feat,
                                                              const float* __restrict__ M_gtv,
                                                              const float* __restrict__ M_glr,
                                                              float* __restrict__ w_gtv, float* __restrict__ w_glr) {
    GLR_SMEM_DECL(smem);
    const int H = s.H, W = s.W, F = s.F, G = s.G, C = G * F;
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const int tile = blockIdx.x % (tiles_w * tiles_h), plane = blockIdx.x / (tiles_w * tiles_h);
    const int g = plane % G, set = (plane / G) % 2, b = plane / (2 * G);
    const int h0 = (tile / tiles_w) * TH, w0 = (tile % tiles_w) * TW;
    const size_t HW = (size_t)H * W;
    constexpr int NH = TH + 2, NW = TW + 2, NP = NH * NW;
    const float* fp = feat + ((size_t)b * 2 * C + (size_t)set * C + (size_t)g * F) * HW;
    const float* Mg = (set ? M_glr : M_gtv) + g * F;
    float* wp = (set ? w_glr : w_gtv) + ((size_t)b * G + g) * 4 * HW;
    // phase 1: normalised, scaled features on the tile (+) 1 (clamp-extended), layout ft[f][pixel]
    for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        int h = glr_clampi(h0 - 1 + i / NW, 0, H - 1), w = glr_clampi(w0 - 1 + i % NW, 0, W - 1);
        const float* q = fp + (size_t)h * W + w;
        float n2 = 0.f;
        for (int f = 0; f < F; ++f) { float v = q[f * HW]; n2 += v * v; }
        float nrm = fmaxf(sqrtf(n2), 1e-12f);
        for (int f = 0; f < F; ++f) smem[f * NP + i] = q[f * HW] / nrm * Mg[f];
    }
    __syncthreads();
    // phase 2: similarities with the four neighbours, softmax
    for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
        int lh = i / TW, lw = i % TW, h = h0 + lh, w = w0 + lw;
        if (h >= H || w >= W) continue;
        const float* c = smem + (lh + 1) * NW + (lw + 1);
        float su = 0.f, sl = 0.f, sr = 0.f, sd = 0.f;
        for (int f = 0; f < F; ++f) {
            const float* cf = c + f * NP;
            float v = cf[0];
            su += v * cf[-NW]; sl += v * cf[-1]; sr += v * cf[1]; sd += v * cf[NW];
        }
        float mx = fmaxf(fmaxf(su, sl), fmaxf(sr, sd));
        su = expf(su - mx); sl = expf(sl - mx); sr = expf(sr - mx); sd = expf(sd - mx);
        float inv = 1.f / (su + sl + sr + sd);
        size_t gi = (size_t)h * W + w;
        wp[gi] = su * inv; wp[HW + gi] = sl * inv; wp[2 * HW + gi] = sr * inv; wp[3 * HW + gi] = sd * inv;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#ifndef GLR_TH
#define GLR_TH 32
#endif
#ifndef GLR_TW
#define GLR_TW 32
#endif
#define GLR_WT_TH 16
#define GLR_WT_TW 32

template <int MODE, bool GEN>
static int launch_stage_gen(const BlockFwdArgs& a, void* stream) {
    const glrgtv_shape& s = a.s;
    const long tiles = (long)((s.W + GLR_TW - 1) / GLR_TW) * ((s.H + GLR_TH - 1) / GLR_TH);
    const long blocks = tiles * s.B * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
#ifndef GLR_FWD_NT
#define GLR_FWD_NT 256
#endif
    constexpr int NT = GLR_FWD_NT;
    constexpr size_t smem = (size_t)FwdLayout<MODE, GLR_TH, GLR_TW, NT, GEN>::total * sizeof(float);
    static_assert(smem <= 227 * 1024, "forward tile does not fit shared memory");
#ifndef GLRGTV_EMU
    static size_t optin[GLR_MAX_DEVICES] = {0};
    if (int rc_ = glr_smem_optin(k_block_stage<MODE, GLR_TH, GLR_TW, NT, GEN>, smem, optin)) return rc_;
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_BA + MODE, stream);
    GLR_LAUNCH((k_block_stage<MODE, GLR_TH, GLR_TW, NT, GEN>), dim3((unsigned)blocks), NT, smem, stream, a);
    GLR_PROF_END(GLRGTV_SLOT_FWD_BA + MODE, stream);
    return GLR_CHECK_LAUNCH();
}
// W % 8 == 0 (quads never straddle the image border at either resolution) takes the branch-free kernels
template <int MODE>
static int launch_stage(const BlockFwdArgs& a, void* stream) {
    return (a.s.W % 8 == 0) ? launch_stage_gen<MODE, false>(a, stream) : launch_stage_gen<MODE, true>(a, stream);
}

int glr_weights_walk_fwd(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, float* wt, float* wl, void* stream);

static int launch_weights(const glrgtv_shape& s, const float* feat, const float* Mt, const float* Ml, float* wt,
                          float* wl, void* stream) {
    {   // the row walkers of weights_walk.cu where the shape is theirs (W % 4 == 0, F = 6 or 12), else the tile kernel below
        GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_WEIGHTS, stream);
        const int rcw = glr_weights_walk_fwd(s, feat, Mt, Ml, wt, wl, stream);
        GLR_PROF_END(GLRGTV_SLOT_FWD_WEIGHTS, stream);
        if (rcw != GLRGTV_ERR_UNSUPPORTED) return rcw;
    }
    const long tiles = (long)((s.W + GLR_WT_TW - 1) / GLR_WT_TW) * ((s.H + GLR_WT_TH - 1) / GLR_WT_TH);
    const long blocks = tiles * s.B * 2 * s.G;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    const size_t smem = (size_t)s.F * (GLR_WT_TH + 2) * (GLR_WT_TW + 2) * sizeof(float);
    if (smem > 200 * 1024) return GLRGTV_ERR_UNSUPPORTED;
#ifndef GLRGTV_EMU
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(k_block_weights<GLR_WT_TH, GLR_WT_TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return glr_record_launch_error();
    }
#endif
    GLR_PROF_BEGIN(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    GLR_LAUNCH((k_block_weights<GLR_WT_TH, GLR_WT_TW>), dim3((unsigned)blocks), GLR_THREADS, smem, stream, s, feat, Mt,
               Ml, wt, wl);
    GLR_PROF_END(GLRGTV_SLOT_FWD_WEIGHTS, stream);
    return GLR_CHECK_LAUNCH();
}

int glr_block_params_ok(const glrgtv_shape* s, const glrgtv_block_params* p) {
    if (!p) return GLRGTV_ERR_POINTER;
    const glrgtv_opparams* ops[4] = {&p->gtv0, &p->glr0, &p->gtv1, &p->glr1};
    for (int i = 0; i < 4; ++i) {
        const glrgtv_stats& st = ops[i]->stats;
        if (!glr_aligned(st.p01) || !glr_aligned(st.p02a) || !glr_aligned(st.p02b) || !glr_aligned(st.p03) ||
            !glr_aligned(ops[i]->multiM))
            return GLRGTV_ERR_POINTER;
        if (st.n != s->G * s->F || st.pad != GLRGTV_PAD_CLAMP) return GLRGTV_ERR_SHAPE;
    }
    const float* v[8] = {p->alpha, p->beta, p->mu0, p->ro0, p->gamma0, p->mu1, p->ro1, p->gamma1};
    for (int i = 0; i < 8; ++i)
        if (!glr_aligned(v[i])) return GLRGTV_ERR_POINTER;
    if (p->skip && !glr_aligned(p->skip)) return GLRGTV_ERR_POINTER;
    return GLRGTV_OK;
}

// block_stream_fwd.cu
extern int g_glr_block_path;
int glr_stream_fwd_eligible(const glrgtv_shape* s);
int glr_launch_gtv_coeffs(const glrgtv_shape& s, const float* w, float* c, void* stream);
int glr_stream_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x, float* out,
                         const glrgtv_block_saved* sv, void* stream);

int glr_stream_block_fwd_stage(int stage, const glrgtv_shape* s, const glrgtv_block_params* p, const float* x, float* out,
                               const glrgtv_block_saved* sv, int row0, int row1, void* stream);

// One piece of the block forward, for callers that interleave the stages with their own work (spatially sharded
// inference: a halo exchange between stages).  stage 0 = edge weights + GTV coefficients of the WHOLE local plane;
// stages 1..4 = BA, X1, X2, X3 on the rows [row0, row1).  Streaming kernels only (W % 8 == 0).
extern "C" int glrgtv_block_fwd_stage(int stage, const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                                      const float* feat0, const float* feat1, float* out, const glrgtv_block_saved* sv,
                                      int row0, int row1, void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    int rc = glr_block_params_ok(s, p);
    if (rc) return rc;
    if (!sv || !glr_stream_fwd_eligible(s)) return GLRGTV_ERR_UNSUPPORTED;
    float* need[11] = {sv->wT0, sv->wL0, sv->wT1, sv->wL1, sv->bA, sv->x1, sv->bB, sv->r1, sv->x2, sv->cT0, sv->cT1};
    for (int i = 0; i < 11; ++i)
        if (!glr_aligned(need[i]) || !glr_aligned16(need[i])) return GLRGTV_ERR_POINTER;
    GLR_REQUIRE_PTR(x);
    if (!glr_aligned16(x)) return GLRGTV_ERR_POINTER;
    if (stage == 0) {
        GLR_REQUIRE_PTR(feat0); GLR_REQUIRE_PTR(feat1);
        glrgtv_shape sc = *s;
        sc.H /= 2; sc.W /= 2;
        if ((rc = launch_weights(*s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, stream))) return rc;
        if ((rc = launch_weights(sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, stream))) return rc;
        if ((rc = glr_launch_gtv_coeffs(*s, sv->wT0, sv->cT0, stream))) return rc;
        return glr_launch_gtv_coeffs(sc, sv->wT1, sv->cT1, stream);
    }
    if (stage < 1 || stage > 4 || row0 < 0 || row1 > s->H || row0 >= row1 || (row0 & 1) || (row1 & 1)) return GLRGTV_ERR_SHAPE;
    if (stage == 4) { GLR_REQUIRE_PTR(out); if (!glr_aligned16(out)) return GLRGTV_ERR_POINTER; }
    return glr_stream_block_fwd_stage(stage - 1, s, p, x, out, sv, row0, row1, stream);
}

extern "C" int glrgtv_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                                const float* feat0, const float* feat1, float* out, const glrgtv_block_saved* sv,
                                void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    int rc = glr_block_params_ok(s, p);
    if (rc) return rc;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(feat0); GLR_REQUIRE_PTR(feat1); GLR_REQUIRE_PTR(out);
    if (!sv) return GLRGTV_ERR_POINTER;
    float* need[9] = {sv->wT0, sv->wL0, sv->wT1, sv->wL1, sv->bA, sv->x1, sv->bB, sv->r1, sv->x2};
    for (int i = 0; i < 9; ++i) GLR_REQUIRE_PTR(need[i]);
    // the float4 paths need 16-byte aligned tensors (true for any allocator; channel offsets keep it when W%4==0)
    if (!glr_aligned16(x) || !glr_aligned16(out)) return GLRGTV_ERR_POINTER;
    for (int i = 0; i < 9; ++i)
        if (!glr_aligned16(need[i])) return GLRGTV_ERR_POINTER;

    glrgtv_shape sc = *s;
    sc.H /= 2; sc.W /= 2;
    if ((rc = launch_weights(*s, feat0, p->gtv0.multiM, p->glr0.multiM, sv->wT0, sv->wL0, stream))) return rc;
    if ((rc = launch_weights(sc, feat1, p->gtv1.multiM, p->glr1.multiM, sv->wT1, sv->wL1, stream))) return rc;

    // register-streaming stage kernels (block_stream_fwd.cu) where the shape allows, else the plane kernels below
    const bool can_stream = glr_stream_fwd_eligible(s) && sv->cT0 && sv->cT1 && glr_aligned16(sv->cT0) && glr_aligned16(sv->cT1);
    if (g_glr_block_path == 2 && !can_stream) return GLRGTV_ERR_UNSUPPORTED;
    // the symmetric GTV coefficients are part of the saved state WHATEVER kernels run the stages: the backward may take the
    // streaming kernels even when this forward took the plane kernels (it picks its path from the shape, not from this call)
    if (sv->cT0 && sv->cT1) {
        if ((rc = glr_launch_gtv_coeffs(*s, sv->wT0, sv->cT0, stream))) return rc;
        if ((rc = glr_launch_gtv_coeffs(sc, sv->wT1, sv->cT1, stream))) return rc;
    }
    if (can_stream && g_glr_block_path != 1) return glr_stream_block_fwd(s, p, x, out, sv, stream);

    BlockFwdArgs a;
    a.s = *s; a.p = *p;
    a.wT0 = sv->wT0; a.wL0 = sv->wL0; a.wT1 = sv->wT1; a.wL1 = sv->wL1;
    a.y = nullptr; a.bB_in = nullptr; a.r1_in = nullptr; a.out1 = nullptr; a.out2 = nullptr;
    a.z = x; a.out0 = sv->bA;
    if ((rc = launch_stage<MODE_BA>(a, stream))) return rc;
    a.z = sv->bA; a.out0 = sv->x1;
    if ((rc = launch_stage<MODE_X1>(a, stream))) return rc;
    a.z = sv->x1; a.y = x; a.out0 = sv->x2; a.out1 = sv->bB; a.out2 = sv->r1;
    if ((rc = launch_stage<MODE_X2>(a, stream))) return rc;
    a.z = sv->x2; a.y = x; a.bB_in = sv->bB; a.r1_in = sv->r1; a.out0 = out; a.out1 = nullptr; a.out2 = nullptr;
    return launch_stage<MODE_X3>(a, stream);
}
