// ops_basic.cu - one kernel per public operator of GLRFast / GTVFast, forward and VJP.
//
// These back the public methods of the drop-in modules (extract_edge_weights, stats_conv,
// stats_conv_transpose, op_L_norm, op_C, op_C_transpose, soft_threshold) for ANY window (V1X0's
// 3x3 cross as well as the older family's 8/12/24-edge windows) and both padding rules of S.
// They read neighbours straight from global memory (L1/L2 serve the re-reads); the fused,
// shared-memory-tiled kernels of the block itself live in block_fwd.cu / block_bwd.cu.
//
// Math: SURVEY.md Appendix B.1-B.8 (derived from V1X0:128-237, 359-523, 684-704).
#include "common.cuh"

#define PIX_LOOP(i, n) for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < (n); i += gridDim.y * blockDim.x)

// ================================================================== B.1 edge weights
// plane = (b,g).  w[e] is used as scratch for the similarities before the softmax.
__global__ void k_edge_weights_fwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ feat,
                                   const float* __restrict__ M, float* __restrict__ w) {
    const int HW = s.H * s.W, F = s.F, E = win.n_edges;
    const int g = blockIdx.x % s.G;
    const float* fp = feat + (size_t)blockIdx.x * F * HW;
    float* wp = w + (size_t)blockIdx.x * E * HW;
    const float* Mg = M + g * F;
    PIX_LOOP(i, HW) {
        int h = i / s.W, x = i % s.W;
        float n0 = 0.f;
        for (int f = 0; f < F; ++f) { float v = fp[f * HW + i]; n0 += v * v; }
        n0 = fmaxf(sqrtf(n0), 1e-12f);
        float mx = -INFINITY;
        for (int e = 0; e < E; ++e) {
            int j = glr_clampi(h + win.dh[e], 0, s.H - 1) * s.W + glr_clampi(x + win.dw[e], 0, s.W - 1);
            float nq = 0.f;
            for (int f = 0; f < F; ++f) { float v = fp[f * HW + j]; nq += v * v; }
            nq = fmaxf(sqrtf(nq), 1e-12f);
            float acc = 0.f;
            for (int f = 0; f < F; ++f) {
                float m = Mg[f];
                acc += (fp[f * HW + i] / n0 * m) * (fp[f * HW + j] / nq * m);
            }
            wp[e * HW + i] = acc;
            mx = fmaxf(mx, acc);
        }
        float sum = 0.f;
        for (int e = 0; e < E; ++e) {
            float v = expf(wp[e * HW + i] - mx);
            wp[e * HW + i] = v;
            sum += v;
        }
        for (int e = 0; e < E; ++e) wp[e * HW + i] = wp[e * HW + i] / sum;
    }
}

// softmax VJP gs_e = w_e (gw_e - sum w gw) -> scratch[0 : E*HW] per plane, raw norm -> scratch tail
__global__ void k_edge_weights_bwd_a(glrgtv_shape s, int E, const float* __restrict__ feat, size_t feat_bs,
                                     const float* __restrict__ w, const float* __restrict__ gw,
                                     float* __restrict__ gs, float* __restrict__ nrm) {
    const int HW = s.H * s.W, F = s.F;
    const float* fp = feat + (size_t)(blockIdx.x / s.G) * feat_bs + (size_t)(blockIdx.x % s.G) * F * HW;
    const float* wp = w + (size_t)blockIdx.x * E * HW;
    const float* gp = gw + (size_t)blockIdx.x * E * HW;
    float* gsp = gs + (size_t)blockIdx.x * E * HW;
    float* np_ = nrm + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        float dot = 0.f;
        for (int e = 0; e < E; ++e) dot += wp[e * HW + i] * gp[e * HW + i];
        for (int e = 0; e < E; ++e) gsp[e * HW + i] = wp[e * HW + i] * (gp[e * HW + i] - dot);
        float n0 = 0.f;
        for (int f = 0; f < F; ++f) { float v = fp[f * HW + i]; n0 += v * v; }
        np_[i] = sqrtf(n0);
    }
}

// gfeat, gM.  plane = (b,g); gridDim.y must be 1 x (pixels handled by the strided loop) so that the
// per-(g,f) partial sums of one block cover its pixels exactly once.
__global__ void k_edge_weights_bwd_b(glrgtv_shape s, glrgtv_window win, const float* __restrict__ feat,
                                     size_t feat_bs, const float* __restrict__ M, const float* __restrict__ gs,
                                     const float* __restrict__ nrm, float* __restrict__ gfeat, size_t gfeat_bs,
                                     float* __restrict__ gM) {
    GLR_SMEM_DECL(red);
    const int H = s.H, W = s.W, HW = H * W, F = s.F, E = win.n_edges;
    const int g = blockIdx.x % s.G;
    const float* fp = feat + (size_t)(blockIdx.x / s.G) * feat_bs + (size_t)g * F * HW;
    const float* gsp = gs + (size_t)blockIdx.x * E * HW;
    const float* np_ = nrm + (size_t)blockIdx.x * HW;
    float* gfp = gfeat + (size_t)(blockIdx.x / s.G) * gfeat_bs + (size_t)g * F * HW;
    const float* Mg = M + g * F;
    // sweep 1: gft_f (gradient wrt the normalised+scaled feature) -> gfeat (temporary), gM partials
    for (int f = 0; f < F; ++f) {
        const float m = Mg[f];
        float part = 0.f;
        PIX_LOOP(i, HW) {
            int h = i / W, x = i % W;
            float acc = 0.f;
            for (int e = 0; e < E; ++e) {
                int dh = win.dh[e], dw = win.dw[e];
                int j = glr_clampi(h + dh, 0, H - 1) * W + glr_clampi(x + dw, 0, W - 1);
                float inv_j = 1.f / fmaxf(np_[j], 1e-12f);
                acc += gsp[e * HW + i] * (fp[f * HW + j] * inv_j * m);
                int hlo, hhi, wlo, whi;
                glr_clamp_preimage(h, dh, H, hlo, hhi);
                glr_clamp_preimage(x, dw, W, wlo, whi);
                for (int ph = hlo; ph <= hhi; ++ph)
                    for (int pw = wlo; pw <= whi; ++pw) {
                        int p = ph * W + pw;
                        float inv_p = 1.f / fmaxf(np_[p], 1e-12f);
                        acc += gsp[e * HW + p] * (fp[f * HW + p] * inv_p * m);
                    }
            }
            gfp[f * HW + i] = acc;
            part += acc * (fp[f * HW + i] / fmaxf(np_[i], 1e-12f));
        }
        float tot = block_sum(part, red);
        if (threadIdx.x == 0) atomicAdd(&gM[g * F + f], tot);
    }
    // sweep 2: through  fhat = f / max(|f|, eps)  (each thread revisits exactly the pixels it wrote)
    PIX_LOOP(i, HW) {
        float nr = np_[i];
        float inv = 1.f / fmaxf(nr, 1e-12f);
        float dot = 0.f;
        if (nr > 1e-12f)
            for (int f = 0; f < F; ++f) dot += (fp[f * HW + i] * inv) * (Mg[f] * gfp[f * HW + i]);
        for (int f = 0; f < F; ++f) {
            float gh = Mg[f] * gfp[f * HW + i];
            gfp[f * HW + i] = (gh - (fp[f * HW + i] * inv) * dot) * inv;
        }
    }
}

// ================================================================== B.2 / B.3 stats conv
// plane = (b,c)
__global__ void k_stats_conv_fwd(glrgtv_shape s, glrgtv_stats st, const float* __restrict__ x,
                                 float* __restrict__ out) {
    const int H = s.H, W = s.W, HW = H * W, C = s.G * s.F;
    const StatsTaps k = glr_load_taps(st, blockIdx.x % C);
    const float* xp = x + (size_t)blockIdx.x * HW;
    float* op = out + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, w = i % W;
        int hu = glr_mapi(h - 1, H, st.pad), hd = glr_mapi(h + 1, H, st.pad);
        int wl = glr_mapi(w - 1, W, st.pad), wr = glr_mapi(w + 1, W, st.pad);
        op[i] = k.kc * xp[i] + k.kr * xp[h * W + wr] + k.kd * xp[hd * W + w] + k.ku * xp[hu * W + w] +
                k.kl * xp[h * W + wl];
    }
}

__device__ __forceinline__ void glr_stats_grad_commit(const glrgtv_stats& st, int c, float* gstats, float gc,
                                                      float gr, float gd, float gu, float gl) {
    int n = st.n, i = n == 1 ? 0 : c;
    atomicAdd(&gstats[0 * n + i], gc);
    atomicAdd(&gstats[1 * n + i], gr - gc);
    atomicAdd(&gstats[2 * n + i], gd - gc);
    atomicAdd(&gstats[3 * n + i], 4.f * gc - gr - gd - gu - gl);
}

__global__ void k_stats_conv_bwd(glrgtv_shape s, glrgtv_stats st, const float* __restrict__ x,
                                 const float* __restrict__ g, float* __restrict__ gx, float* __restrict__ gstats) {
    GLR_SMEM_DECL(red);
    const int H = s.H, W = s.W, HW = H * W, C = s.G * s.F, c = blockIdx.x % C;
    const StatsTaps k = glr_load_taps(st, c);
    const float* xp = x + (size_t)blockIdx.x * HW;
    const float* gp = g + (size_t)blockIdx.x * HW;
    float* gxp = gx + (size_t)blockIdx.x * HW;
    float sc = 0.f, sr = 0.f, sd = 0.f, su = 0.f, sl = 0.f;
    PIX_LOOP(i, HW) {
        int h = i / W, w = i % W;
        int hu = glr_mapi(h - 1, H, st.pad), hd = glr_mapi(h + 1, H, st.pad);
        int wl = glr_mapi(w - 1, W, st.pad), wr = glr_mapi(w + 1, W, st.pad);
        float gi = gp[i];
        sc += gi * xp[i];
        sr += gi * xp[h * W + wr];
        sd += gi * xp[hd * W + w];
        su += gi * xp[hu * W + w];
        sl += gi * xp[h * W + wl];
        // gx[q] = sum_t k_t sum_{p : map(p+o_t) = q} g[p]; candidates are within 2 pixels along one axis
        float acc = k.kc * gi;
        for (int pw = w - 2; pw <= w + 2; ++pw) {
            if (pw < 0 || pw >= W) continue;
            if (glr_mapi(pw + 1, W, st.pad) == w) acc += k.kr * gp[h * W + pw];
            if (glr_mapi(pw - 1, W, st.pad) == w) acc += k.kl * gp[h * W + pw];
        }
        for (int ph = h - 2; ph <= h + 2; ++ph) {
            if (ph < 0 || ph >= H) continue;
            if (glr_mapi(ph + 1, H, st.pad) == h) acc += k.kd * gp[ph * W + w];
            if (glr_mapi(ph - 1, H, st.pad) == h) acc += k.ku * gp[ph * W + w];
        }
        gxp[i] = acc;
    }
    sc = block_sum(sc, red); sr = block_sum(sr, red); sd = block_sum(sd, red);
    su = block_sum(su, red); sl = block_sum(sl, red);
    if (threadIdx.x == 0) glr_stats_grad_commit(st, c, gstats, sc, sr, sd, su, sl);
}

// St y[q] = sum_t k_t y[q-o_t] [inside]
__global__ void k_stats_conv_t_fwd(glrgtv_shape s, glrgtv_stats st, const float* __restrict__ y,
                                   float* __restrict__ out) {
    const int H = s.H, W = s.W, HW = H * W, C = s.G * s.F;
    const StatsTaps k = glr_load_taps(st, blockIdx.x % C);
    const float* yp = y + (size_t)blockIdx.x * HW;
    float* op = out + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, w = i % W;
        float acc = k.kc * yp[i];
        if (w - 1 >= 0) acc += k.kr * yp[i - 1];
        if (h - 1 >= 0) acc += k.kd * yp[i - W];
        if (h + 1 < H) acc += k.ku * yp[i + W];
        if (w + 1 < W) acc += k.kl * yp[i + 1];
        op[i] = acc;
    }
}

__global__ void k_stats_conv_t_bwd(glrgtv_shape s, glrgtv_stats st, const float* __restrict__ y,
                                   const float* __restrict__ g, float* __restrict__ gy, float* __restrict__ gstats) {
    GLR_SMEM_DECL(red);
    const int H = s.H, W = s.W, HW = H * W, C = s.G * s.F, c = blockIdx.x % C;
    const StatsTaps k = glr_load_taps(st, c);
    const float* yp = y + (size_t)blockIdx.x * HW;
    const float* gp = g + (size_t)blockIdx.x * HW;
    float* gyp = gy + (size_t)blockIdx.x * HW;
    float sc = 0.f, sr = 0.f, sd = 0.f, su = 0.f, sl = 0.f;
    PIX_LOOP(i, HW) {
        int h = i / W, w = i % W;
        float gi = gp[i];
        // gy[p] = sum_t k_t g[p+o_t] [inside]
        float acc = k.kc * gi;
        if (w + 1 < W) acc += k.kr * gp[i + 1];
        if (h + 1 < H) acc += k.kd * gp[i + W];
        if (h - 1 >= 0) acc += k.ku * gp[i - W];
        if (w - 1 >= 0) acc += k.kl * gp[i - 1];
        gyp[i] = acc;
        // gk_t = sum_q g[q] y[q-o_t] [inside]
        sc += gi * yp[i];
        if (w - 1 >= 0) sr += gi * yp[i - 1];
        if (h - 1 >= 0) sd += gi * yp[i - W];
        if (h + 1 < H) su += gi * yp[i + W];
        if (w + 1 < W) sl += gi * yp[i + 1];
    }
    sc = block_sum(sc, red); sr = block_sum(sr, red); sd = block_sum(sd, red);
    su = block_sum(su, red); sl = block_sum(sl, red);
    if (threadIdx.x == 0) glr_stats_grad_commit(st, c, gstats, sc, sr, sd, su, sl);
}

// ================================================================== B.4 L
// plane = (b,c)
__global__ void k_op_L_fwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ x,
                           const float* __restrict__ w, float* __restrict__ out) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* xp = x + (size_t)blockIdx.x * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* op = out + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float acc = 0.f;
        for (int e = 0; e < E; ++e) {
            int j = glr_clampi(h + win.dh[e], 0, H - 1) * W + glr_clampi(c + win.dw[e], 0, W - 1);
            acc += wp[e * HW + i] * xp[j];
        }
        op[i] = xp[i] - acc;
    }
}
// gx: plane = (b,c)
__global__ void k_op_L_bwd_x(glrgtv_shape s, glrgtv_window win, const float* __restrict__ w,
                             const float* __restrict__ g, float* __restrict__ gx) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* gp = g + (size_t)blockIdx.x * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* gxp = gx + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float acc = 0.f;
        for (int e = 0; e < E; ++e) {
            int hlo, hhi, wlo, whi;
            glr_clamp_preimage(h, win.dh[e], H, hlo, hhi);
            glr_clamp_preimage(c, win.dw[e], W, wlo, whi);
            for (int ph = hlo; ph <= hhi; ++ph)
                for (int pw = wlo; pw <= whi; ++pw) acc += wp[e * HW + ph * W + pw] * gp[ph * W + pw];
        }
        gxp[i] = gp[i] - acc;
    }
}
// gw: plane = (b,g);  sign = -1 for L (gw_e = -sum_f g_f x_f[n_e]).
__global__ void k_op_L_bwd_w(glrgtv_shape s, glrgtv_window win, const float* __restrict__ x,
                             const float* __restrict__ g, float* __restrict__ gw) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges, F = s.F;
    const float* xp = x + (size_t)blockIdx.x * F * HW;
    const float* gp = g + (size_t)blockIdx.x * F * HW;
    float* gwp = gw + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        for (int e = 0; e < E; ++e) {
            int j = glr_clampi(h + win.dh[e], 0, H - 1) * W + glr_clampi(c + win.dw[e], 0, W - 1);
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc += gp[f * HW + i] * xp[f * HW + j];
            gwp[e * HW + i] = -acc;
        }
    }
}

// ================================================================== B.5 C (after S)
// plane = (b,c); z [B,G,F,E,H,W]
__global__ void k_op_C_fwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ sx,
                           const float* __restrict__ w, float* __restrict__ z) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* sp = sx + (size_t)blockIdx.x * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* zp = z + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float v = sp[i];
        for (int e = 0; e < E; ++e) {
            int j = glr_clampi(h + win.dh[e], 0, H - 1) * W + glr_clampi(c + win.dw[e], 0, W - 1);
            zp[e * HW + i] = wp[e * HW + i] * (v - sp[j]);
        }
    }
}
__global__ void k_op_C_bwd_x(glrgtv_shape s, glrgtv_window win, const float* __restrict__ w,
                             const float* __restrict__ gz, float* __restrict__ gsx) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* gp = gz + (size_t)blockIdx.x * E * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* op = gsx + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float acc = 0.f;
        for (int e = 0; e < E; ++e) {
            acc += wp[e * HW + i] * gp[e * HW + i];
            int hlo, hhi, wlo, whi;
            glr_clamp_preimage(h, win.dh[e], H, hlo, hhi);
            glr_clamp_preimage(c, win.dw[e], W, wlo, whi);
            for (int ph = hlo; ph <= hhi; ++ph)
                for (int pw = wlo; pw <= whi; ++pw) {
                    int p = ph * W + pw;
                    acc -= wp[e * HW + p] * gp[e * HW + p];
                }
        }
        op[i] = acc;
    }
}
// plane = (b,g)
__global__ void k_op_C_bwd_w(glrgtv_shape s, glrgtv_window win, const float* __restrict__ sx,
                             const float* __restrict__ gz, float* __restrict__ gw) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges, F = s.F;
    const float* sp = sx + (size_t)blockIdx.x * F * HW;
    const float* gp = gz + (size_t)blockIdx.x * F * E * HW;
    float* gwp = gw + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        for (int e = 0; e < E; ++e) {
            int j = glr_clampi(h + win.dh[e], 0, H - 1) * W + glr_clampi(c + win.dw[e], 0, W - 1);
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc += gp[(f * E + e) * HW + i] * (sp[f * HW + i] - sp[f * HW + j]);
            gwp[e * HW + i] = acc;
        }
    }
}

// ================================================================== B.6 Ct (before St)
// plane = (b,c)
__global__ void k_op_Ct_fwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ z,
                            const float* __restrict__ w, float* __restrict__ o) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* zp = z + (size_t)blockIdx.x * E * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* op = o + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float acc = 0.f;
        for (int e = 0; e < E; ++e) {
            acc += wp[e * HW + i] * zp[e * HW + i];
            int ph = h - win.dh[e], pw = c - win.dw[e];
            if (glr_inside(ph, pw, H, W)) acc -= wp[e * HW + ph * W + pw] * zp[e * HW + ph * W + pw];
        }
        op[i] = acc;
    }
}
__global__ void k_op_Ct_bwd_z(glrgtv_shape s, glrgtv_window win, const float* __restrict__ w,
                              const float* __restrict__ go, float* __restrict__ gz) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* gp = go + (size_t)blockIdx.x * HW;
    const float* wp = w + (size_t)(blockIdx.x / s.F) * E * HW;
    float* gzp = gz + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        for (int e = 0; e < E; ++e) {
            int qh = h + win.dh[e], qw = c + win.dw[e];
            float gu = gp[i] - (glr_inside(qh, qw, H, W) ? gp[qh * W + qw] : 0.f);
            gzp[e * HW + i] = wp[e * HW + i] * gu;
        }
    }
}
__global__ void k_op_Ct_bwd_w(glrgtv_shape s, glrgtv_window win, const float* __restrict__ z,
                              const float* __restrict__ go, float* __restrict__ gw) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges, F = s.F;
    const float* zp = z + (size_t)blockIdx.x * F * E * HW;
    const float* gp = go + (size_t)blockIdx.x * F * HW;
    float* gwp = gw + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        for (int e = 0; e < E; ++e) {
            int qh = h + win.dh[e], qw = c + win.dw[e];
            bool in = glr_inside(qh, qw, H, W);
            float acc = 0.f;
            for (int f = 0; f < F; ++f) {
                float gu = gp[f * HW + i] - (in ? gp[f * HW + qh * W + qw] : 0.f);
                acc += zp[(f * E + e) * HW + i] * gu;
            }
            gwp[e * HW + i] = acc;
        }
    }
}

// ================================================================== B.7 soft threshold
// plane = (b,g); n = F*E*H*W contiguous elements per plane
__global__ void k_soft_fwd(int G, int n, const float* __restrict__ t, const float* __restrict__ thr,
                           float* __restrict__ out) {
    const float th = thr[blockIdx.x % G];
    const float* tp = t + (size_t)blockIdx.x * n;
    float* op = out + (size_t)blockIdx.x * n;
    PIX_LOOP(i, n) {
        float v = tp[i];
        op[i] = (v < -th ? v + th : 0.f) + (v > th ? v - th : 0.f);
    }
}
__global__ void k_soft_bwd(int G, int n, const float* __restrict__ t, const float* __restrict__ thr,
                           const float* __restrict__ g, float* __restrict__ gt, float* __restrict__ gthr) {
    GLR_SMEM_DECL(red);
    const int gi = blockIdx.x % G;
    const float th = thr[gi];
    const float* tp = t + (size_t)blockIdx.x * n;
    const float* gp = g + (size_t)blockIdx.x * n;
    float* op = gt + (size_t)blockIdx.x * n;
    float part = 0.f;
    PIX_LOOP(i, n) {
        float v = tp[i], gv = gp[i];
        bool lo = v < -th, hi = v > th;
        op[i] = (lo || hi) ? gv : 0.f;
        part += lo ? gv : (hi ? -gv : 0.f);
    }
    part = block_sum(part, red);
    if (threadIdx.x == 0) atomicAdd(&gthr[gi], part);
}

// ================================================================== B.8 pooling
// plane = (b,c); coarse pixel loop
__global__ void k_pool2(int H, int W, const float* __restrict__ fine, float* __restrict__ coarse) {
    const int Hc = H / 2, Wc = W / 2;
    const float* fp = fine + (size_t)blockIdx.x * H * W;
    float* cp = coarse + (size_t)blockIdx.x * Hc * Wc;
    PIX_LOOP(i, Hc * Wc) {
        int h = i / Wc, w = i % Wc;
        const float* r0 = fp + (2 * h) * W + 2 * w;
        cp[i] = 0.25f * (r0[0] + r0[1] + r0[W] + r0[W + 1]);
    }
}
__global__ void k_unpool2(int H, int W, const float* __restrict__ coarse, float* __restrict__ fine) {
    const int Wc = W / 2;
    const float* cp = coarse + (size_t)blockIdx.x * (H / 2) * Wc;
    float* fp = fine + (size_t)blockIdx.x * H * W;
    PIX_LOOP(i, H * W) {
        int h = i / W, w = i % W;
        fp[i] = 0.25f * cp[(h / 2) * Wc + (w / 2)];
    }
}

// ================================================================== a2 / a3 helpers (public methods)
// normalize_and_transform_features (V1X0:146-157): out = M * f / max(|f|_F, 1e-12).  plane = (b,g)
__global__ void k_normalize_fwd(glrgtv_shape s, const float* __restrict__ feat, const float* __restrict__ M,
                                float* __restrict__ out) {
    const int HW = s.H * s.W, F = s.F;
    const float* fp = feat + (size_t)blockIdx.x * F * HW;
    float* op = out + (size_t)blockIdx.x * F * HW;
    const float* Mg = M + (blockIdx.x % s.G) * F;
    PIX_LOOP(i, HW) {
        float n2 = 0.f;
        for (int f = 0; f < F; ++f) { float v = fp[f * HW + i]; n2 += v * v; }
        float nrm = fmaxf(sqrtf(n2), 1e-12f);
        for (int f = 0; f < F; ++f) op[f * HW + i] = fp[f * HW + i] / nrm * Mg[f];
    }
}
__global__ void k_normalize_bwd(glrgtv_shape s, const float* __restrict__ feat, const float* __restrict__ M,
                                const float* __restrict__ g, float* __restrict__ gfeat, float* __restrict__ gM) {
    GLR_SMEM_DECL(red);
    const int HW = s.H * s.W, F = s.F, gi = blockIdx.x % s.G;
    const float* fp = feat + (size_t)blockIdx.x * F * HW;
    const float* gp = g + (size_t)blockIdx.x * F * HW;
    float* gfp = gfeat + (size_t)blockIdx.x * F * HW;
    const float* Mg = M + gi * F;
    for (int f = 0; f < F; ++f) {
        float part = 0.f;
        PIX_LOOP(i, HW) {
            float n2 = 0.f;
            for (int k = 0; k < F; ++k) { float v = fp[k * HW + i]; n2 += v * v; }
            part += gp[f * HW + i] * (fp[f * HW + i] / fmaxf(sqrtf(n2), 1e-12f));
        }
        part = block_sum(part, red);
        if (threadIdx.x == 0) atomicAdd(&gM[gi * F + f], part);
    }
    PIX_LOOP(i, HW) {
        float n2 = 0.f;
        for (int k = 0; k < F; ++k) { float v = fp[k * HW + i]; n2 += v * v; }
        float nr = sqrtf(n2), inv = 1.f / fmaxf(nr, 1e-12f), dot = 0.f;
        if (nr > 1e-12f)
            for (int k = 0; k < F; ++k) dot += (fp[k * HW + i] * inv) * (Mg[k] * gp[k * HW + i]);
        for (int k = 0; k < F; ++k) gfp[k * HW + i] = (Mg[k] * gp[k * HW + i] - (fp[k * HW + i] * inv) * dot) * inv;
    }
}
// get_neighbors_pixels (V1X0:128-144): out[b,c,e,p] = x[b,c,cl(p+d_e)].  plane = (b,c)
__global__ void k_gather_fwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ x, float* __restrict__ out) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* xp = x + (size_t)blockIdx.x * HW;
    float* op = out + (size_t)blockIdx.x * E * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        for (int e = 0; e < E; ++e)
            op[e * HW + i] = xp[glr_clampi(h + win.dh[e], 0, H - 1) * W + glr_clampi(c + win.dw[e], 0, W - 1)];
    }
}
__global__ void k_gather_bwd(glrgtv_shape s, glrgtv_window win, const float* __restrict__ g, float* __restrict__ gx) {
    const int H = s.H, W = s.W, HW = H * W, E = win.n_edges;
    const float* gp = g + (size_t)blockIdx.x * E * HW;
    float* op = gx + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        int h = i / W, c = i % W;
        float acc = 0.f;
        for (int e = 0; e < E; ++e) {
            int hlo, hhi, wlo, whi;
            glr_clamp_preimage(h, win.dh[e], H, hlo, hhi);
            glr_clamp_preimage(c, win.dw[e], W, wlo, whi);
            for (int ph = hlo; ph <= hhi; ++ph)
                for (int pw = wlo; pw <= whi; ++pw) acc += gp[e * HW + ph * W + pw];
        }
        op[i] = acc;
    }
}

// ================================================================== mixture weighting (V7:1011-1014)
// out[b,c,p] = sum_g x[b,g,c,p] * score[b,g,p].   plane = (b,c)
__global__ void k_mixture_fwd(glrgtv_shape s, const float* __restrict__ x, const float* __restrict__ score,
                              float* __restrict__ out) {
    const int HW = s.H * s.W, G = s.G, F = s.F;
    const int b = blockIdx.x / F, c = blockIdx.x % F;
    const float* xp = x + ((size_t)b * G * F + c) * HW;
    const float* sp = score + (size_t)b * G * HW;
    float* op = out + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        float acc = 0.f;
        for (int g = 0; g < G; ++g) acc += xp[(size_t)g * F * HW + i] * sp[g * HW + i];
        op[i] = acc;
    }
}
// gx[b,g,c,p] = gout[b,c,p] * score[b,g,p];  gscore[b,g,p] = sum_c gout[b,c,p] * x[b,g,c,p].   plane = (b,g)
__global__ void k_mixture_bwd(glrgtv_shape s, const float* __restrict__ x, const float* __restrict__ score,
                              const float* __restrict__ gout, float* __restrict__ gx, float* __restrict__ gscore) {
    const int HW = s.H * s.W, G = s.G, F = s.F;
    const int b = blockIdx.x / G;
    const float* xp = x + (size_t)blockIdx.x * F * HW;
    const float* sp = score + (size_t)blockIdx.x * HW;
    const float* gp = gout + (size_t)b * F * HW;
    float* gxp = gx + (size_t)blockIdx.x * F * HW;
    float* gsp = gscore + (size_t)blockIdx.x * HW;
    PIX_LOOP(i, HW) {
        const float sc = sp[i];
        float acc = 0.f;
        for (int c = 0; c < F; ++c) {
            const float gv = gp[c * HW + i];
            gxp[c * HW + i] = gv * sc;
            acc += gv * xp[c * HW + i];
        }
        gsp[i] = acc;
    }
}

// ================================================================== C ABI
#define PLANES_C(s) ((long)(s)->B * (s)->G * (s)->F)
#define PLANES_G(s) ((long)(s)->B * (s)->G)
#define HW_(s) ((long)(s)->H * (s)->W)

int glr_edge_weights_bwd_strided(const glrgtv_shape* s, const glrgtv_window* win, const float* feat, size_t feat_bs,
                                 const float* multiM, const float* w, const float* gw, float* gfeat, size_t gfeat_bs,
                                 float* gmultiM, float* scratch, void* stream);

extern "C" {

int glrgtv_edge_weights_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* feat,
                            const float* multiM, float* w, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(feat); GLR_REQUIRE_PTR(multiM); GLR_REQUIRE_PTR(w);
    GLR_LAUNCH(k_edge_weights_fwd, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream,
               *s, *win, feat, multiM, w);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_edge_weights_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* feat,
                            const float* multiM, const float* w, const float* gw, float* gfeat,
                            float* gmultiM, float* scratch, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    const size_t bs = (size_t)s->G * s->F * s->H * s->W;
    return glr_edge_weights_bwd_strided(s, win, feat, bs, multiM, w, gw, gfeat, bs, gmultiM, scratch, stream);
}
}  // extern "C"

// feat / gfeat may be channel slices of a wider tensor: feat_bs / gfeat_bs are their batch strides in floats
int glr_edge_weights_bwd_strided(const glrgtv_shape* s, const glrgtv_window* win, const float* feat, size_t feat_bs,
                                 const float* multiM, const float* w, const float* gw, float* gfeat, size_t gfeat_bs,
                                 float* gmultiM, float* scratch, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(feat); GLR_REQUIRE_PTR(multiM); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(gw);
    GLR_REQUIRE_PTR(gfeat); GLR_REQUIRE_PTR(gmultiM); GLR_REQUIRE_PTR(scratch);
    float* gs = scratch;
    float* nrm = scratch + (size_t)PLANES_G(s) * win->n_edges * HW_(s);
    GLR_LAUNCH(k_edge_weights_bwd_a, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream,
               *s, win->n_edges, feat, feat_bs, w, gw, gs, nrm);
    GLR_LAUNCH(k_edge_weights_bwd_b, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 32 * sizeof(float),
               stream, *s, *win, feat, feat_bs, multiM, gs, nrm, gfeat, gfeat_bs, gmultiM);
    return GLR_CHECK_LAUNCH();
}
extern "C" {

static int stats_ok(const glrgtv_shape* s, const glrgtv_stats* st) {
    if (!st || !glr_aligned(st->p01) || !glr_aligned(st->p02a) || !glr_aligned(st->p02b) || !glr_aligned(st->p03))
        return GLRGTV_ERR_POINTER;
    if (st->n != 1 && st->n != s->G * s->F) return GLRGTV_ERR_SHAPE;
    if (st->pad != GLRGTV_PAD_CLAMP && st->pad != GLRGTV_PAD_REFLECT) return GLRGTV_ERR_SHAPE;
    if (st->pad == GLRGTV_PAD_REFLECT && (s->H < 2 || s->W < 2)) return GLRGTV_ERR_SHAPE;
    return GLRGTV_OK;
}

int glrgtv_stats_conv_fwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* x, float* out, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    int rc = stats_ok(s, st);
    if (rc) return rc;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_stats_conv_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *st, x, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_stats_conv_bwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* x, const float* g,
                          float* gx, float* gstats, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    int rc = stats_ok(s, st);
    if (rc) return rc;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gx); GLR_REQUIRE_PTR(gstats);
    GLR_LAUNCH(k_stats_conv_bwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 32 * sizeof(float), stream,
               *s, *st, x, g, gx, gstats);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_stats_conv_t_fwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* y, float* out, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    int rc = stats_ok(s, st);
    if (rc) return rc;
    GLR_REQUIRE_PTR(y); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_stats_conv_t_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *st, y, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_stats_conv_t_bwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* y, const float* g,
                            float* gy, float* gstats, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    int rc = stats_ok(s, st);
    if (rc) return rc;
    GLR_REQUIRE_PTR(y); GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gy); GLR_REQUIRE_PTR(gstats);
    GLR_LAUNCH(k_stats_conv_t_bwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 32 * sizeof(float), stream,
               *s, *st, y, g, gy, gstats);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_op_L_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, const float* w,
                    float* out, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_op_L_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, x, w, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_op_L_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, const float* w,
                    const float* g, float* gx, float* gw, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gx); GLR_REQUIRE_PTR(gw);
    GLR_LAUNCH(k_op_L_bwd_x, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, w, g, gx);
    GLR_LAUNCH(k_op_L_bwd_w, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, x, g, gw);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_op_C_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* sx, const float* w,
                    float* z, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(sx); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(z);
    GLR_LAUNCH(k_op_C_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, sx, w, z);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_op_C_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* sx, const float* w,
                    const float* gz, float* gsx, float* gw, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(sx); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(gz); GLR_REQUIRE_PTR(gsx); GLR_REQUIRE_PTR(gw);
    GLR_LAUNCH(k_op_C_bwd_x, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, w, gz, gsx);
    GLR_LAUNCH(k_op_C_bwd_w, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, sx, gz, gw);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_op_Ct_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* z, const float* w,
                     float* o, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(z); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(o);
    GLR_LAUNCH(k_op_Ct_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, z, w, o);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_op_Ct_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* z, const float* w,
                     const float* go, float* gz, float* gw, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(z); GLR_REQUIRE_PTR(w); GLR_REQUIRE_PTR(go); GLR_REQUIRE_PTR(gz); GLR_REQUIRE_PTR(gw);
    GLR_LAUNCH(k_op_Ct_bwd_z, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, w, go, gz);
    GLR_LAUNCH(k_op_Ct_bwd_w, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, z, go, gw);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_soft_threshold_fwd(const glrgtv_shape* s, int n_edges, const float* t, const float* thr,
                              float* out, void* stream) {
    if (!glr_shape_ok(s) || n_edges <= 0) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(t); GLR_REQUIRE_PTR(thr); GLR_REQUIRE_PTR(out);
    long n = (long)s->F * n_edges * HW_(s);
    if (n > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    GLR_LAUNCH(k_soft_fwd, glr_grid(PLANES_G(s), n, GLR_THREADS * 4), GLR_THREADS, 0, stream, s->G, (int)n, t, thr, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_soft_threshold_bwd(const glrgtv_shape* s, int n_edges, const float* t, const float* thr,
                              const float* g, float* gt, float* gthr, void* stream) {
    if (!glr_shape_ok(s) || n_edges <= 0) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(t); GLR_REQUIRE_PTR(thr); GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gt); GLR_REQUIRE_PTR(gthr);
    long n = (long)s->F * n_edges * HW_(s);
    if (n > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    GLR_LAUNCH(k_soft_bwd, glr_grid(PLANES_G(s), n, GLR_THREADS * 8), GLR_THREADS, 32 * sizeof(float), stream,
               s->G, (int)n, t, thr, g, gt, gthr);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_pool2_fwd(const glrgtv_shape* s, const float* fine, float* coarse, void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(fine); GLR_REQUIRE_PTR(coarse);
    GLR_LAUNCH(k_pool2, glr_grid(PLANES_C(s), HW_(s) / 4, GLR_THREADS), GLR_THREADS, 0, stream, s->H, s->W, fine, coarse);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_unpool2_fwd(const glrgtv_shape* s, const float* coarse, float* fine, void* stream) {
    if (!glr_shape_ok(s) || (s->H & 1) || (s->W & 1)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(coarse); GLR_REQUIRE_PTR(fine);
    GLR_LAUNCH(k_unpool2, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, s->H, s->W, coarse, fine);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_normalize_fwd(const glrgtv_shape* s, const float* feat, const float* multiM, float* out, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(feat); GLR_REQUIRE_PTR(multiM); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_normalize_fwd, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, feat, multiM, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_normalize_bwd(const glrgtv_shape* s, const float* feat, const float* multiM, const float* g,
                         float* gfeat, float* gmultiM, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(feat); GLR_REQUIRE_PTR(multiM); GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gfeat); GLR_REQUIRE_PTR(gmultiM);
    GLR_LAUNCH(k_normalize_bwd, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 32 * sizeof(float), stream,
               *s, feat, multiM, g, gfeat, gmultiM);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_gather_neighbors_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, float* out, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_gather_fwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, x, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_gather_neighbors_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* g, float* gx, void* stream) {
    if (!glr_shape_ok(s) || !glr_window_ok(win)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(g); GLR_REQUIRE_PTR(gx);
    GLR_LAUNCH(k_gather_bwd, glr_grid(PLANES_C(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, *win, g, gx);
    return GLR_CHECK_LAUNCH();
}

int glrgtv_mixture_fwd(const glrgtv_shape* s, const float* x, const float* score, float* out, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(score); GLR_REQUIRE_PTR(out);
    GLR_LAUNCH(k_mixture_fwd, glr_grid((long)s->B * s->F, HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, x, score, out);
    return GLR_CHECK_LAUNCH();
}
int glrgtv_mixture_bwd(const glrgtv_shape* s, const float* x, const float* score, const float* gout, float* gx,
                       float* gscore, void* stream) {
    if (!glr_shape_ok(s)) return GLRGTV_ERR_SHAPE;
    GLR_REQUIRE_PTR(x); GLR_REQUIRE_PTR(score); GLR_REQUIRE_PTR(gout); GLR_REQUIRE_PTR(gx); GLR_REQUIRE_PTR(gscore);
    GLR_LAUNCH(k_mixture_bwd, glr_grid(PLANES_G(s), HW_(s), GLR_THREADS), GLR_THREADS, 0, stream, *s, x, score, gout, gx, gscore);
    return GLR_CHECK_LAUNCH();
}

}  // extern "C"
