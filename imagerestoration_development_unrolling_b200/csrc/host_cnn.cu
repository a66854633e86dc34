// host_cnn.cu - the HBM-bound part of LocalNonLinearBlock (V1X0:911-964), inference forward (first cut).
//
//   reference:  n = CustomLayerNorm(x) = w_n * x / sqrt(var_c(x) + 1e-5)            (V1X0:911-925, per pixel, per sub-net)
//               h = conv1x1(n)  [2 Hd channels]                                      (V1X0:933-937)
//               m = depthwise 3x3 (replicate padded) of h; gate, val = m.chunk(2)    (V1X0:938-946)
//               out = s0 x + s1 conv1x1(sigmoid(gate) gate val)                      (V1X0:947, 961-964)
//
// The per-pixel scale rs = 1/sqrt(var + eps) commutes with the 1x1 convolution, so the normalised tensor is never written:
//   k_pixel_rstd :  x -> rs [B, nsub, H, W]                                  reads x once (4 C bytes per pixel)
//   (GEMM, cuBLAS):  h_raw = (W1 diag(w_n)) x
//   k_dwconv_gate:  h_raw, rs -> u = sigmoid(G) G V,  G | V = dw3x3(rs h_raw)   reads h once, writes u (12 Hd bytes per pixel)
//   (GEMM, cuBLAS):  out = s0 x + (s1 W2) u
// instead of the reference's eleven passes (variance, divide, scale, pad, conv, chunk copies, sigmoid, two products, skip).
// On a row strip of a spatially sharded image the rows above / below come from the neighbours as already scaled rows
// (`top` / `bot`, [B, 2Hd, W]); NULL means the true image border (replicate).
#include "tile.cuh"

// one item = 4 consecutive pixels of one (image, sub-net).  Two passes over the sub-net's channels, as torch.var does: the mean first,
// then the squared deviations from it (an error of the rounded mean enters the sum only to second order; a one-pass Welford update
// is first-order sensitive to it and loses 1e-5 when the spread is small against the mean; a sum / sum-of-squares form cancels).
// The second pass re-reads x from L1 / L2.
__global__ void __launch_bounds__(256) k_pixel_rstd(const float* __restrict__ x, float* __restrict__ rs, int B, int nsub, int c,
                                                    long HW, float eps) {
    const long Q = HW / 4, total = (long)B * nsub * Q;
    const float invc = 1.f / (float)c, invn = 1.f / (float)(c - 1);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long q = i % Q, bs = i / Q;
        const float* p = x + bs * c * HW + 4 * q;
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f}, v0[4];
        ld4(p, v0);
        for (int k = 1; k < c; ++k) {                 // mean = x_0 + mean(x - x_0): the differences are small, so their sum rounds little
            float v[4];
            ld4(p + (long)k * HW, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) mean[j] += v[j] - v0[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) mean[j] = v0[j] + mean[j] * invc;
        for (int k = 0; k < c; ++k) {
            float v[4];
            ld4(p + (long)k * HW, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float d = v[j] - mean[j]; m2[j] += d * d; }
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = 1.f / sqrtf(m2[j] * invn + eps);
        st4(rs + bs * HW + 4 * q, o);
    }
}

// a scaled row segment: the thread's 4 columns and the scalars left / right of them
struct SegRow {
    float l, v[4], r;
};
#define DW_BAND 32      // rows one thread walks (2 halo rows re-read per band: 6 %)

struct DwArgs {
    const float *h, *rs, *w9, *top, *bot;
    float* u;
    int B, Hd, nsub, H, W, n_bands;
};

// raw operands of one row of a gate / value channel pair: the thread's quads of h and of rs, and - on the lanes at a warp's
// edge only - the already scaled scalars left / right of the quad (the other lanes get them from their neighbours' registers)
struct RawRow {
    float g[4], v[4], sg[4], sv[4], gl, gr, vl, vr;
};
template <bool ONE>      // ONE: a single sub-net, gate and value share rs
__device__ __forceinline__ RawRow dw_load(const DwArgs& a, int b, int k, int row, int q, bool need_l, bool need_r) {
    RawRow o;
    const int cg = k, cv = a.Hd + k;
    if ((row < 0 && a.top) || (row >= a.H && a.bot)) {           // the neighbour strip's row: already scaled
        const float* base = (row < 0 ? a.top : a.bot) + (long)b * 2 * a.Hd * a.W + 4 * q;
        const float *pg = base + (long)cg * a.W, *pv = base + (long)cv * a.W;
        ld4(pg, o.g); ld4(pv, o.v);
#pragma unroll
        for (int j = 0; j < 4; ++j) o.sg[j] = o.sv[j] = 1.f;
        o.gl = need_l ? pg[-1] : 0.f; o.vl = need_l ? pv[-1] : 0.f;
        o.gr = need_r ? pg[4] : 0.f; o.vr = need_r ? pv[4] : 0.f;
        return o;
    }
    row = row < 0 ? 0 : row >= a.H ? a.H - 1 : row;              // the true image border replicates
    const float* pg = a.h + (((long)b * 2 * a.Hd + cg) * a.H + row) * a.W + 4 * q;
    const float* pv = a.h + (((long)b * 2 * a.Hd + cv) * a.H + row) * a.W + 4 * q;
    const int per = 2 * a.Hd / a.nsub;
    const float* rg = a.rs + (((long)b * a.nsub + (ONE ? 0 : cg / per)) * a.H + row) * a.W + 4 * q;
    const float* rv = ONE ? rg : a.rs + (((long)b * a.nsub + cv / per) * a.H + row) * a.W + 4 * q;
    ld4(pg, o.g); ld4(pv, o.v); ld4(rg, o.sg);
    if (!ONE) ld4(rv, o.sv);
    o.gl = need_l ? pg[-1] * rg[-1] : 0.f; o.vl = need_l ? pv[-1] * rv[-1] : 0.f;
    o.gr = need_r ? pg[4] * rg[4] : 0.f; o.vr = need_r ? pv[4] * rv[4] : 0.f;
    return o;
}
// scale, and fetch the scalars beside the quad from the neighbouring lanes (whole warp calls this)
template <bool ONE>
__device__ __forceinline__ void dw_finish(const RawRow& w, bool first, bool last, bool need_l, bool need_r, SegRow& G, SegRow& V) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { G.v[j] = w.g[j] * w.sg[j]; V.v[j] = w.v[j] * (ONE ? w.sg[j] : w.sv[j]); }
    const float gl = __shfl_up_sync(0xffffffffu, G.v[3], 1), gr = __shfl_down_sync(0xffffffffu, G.v[0], 1);
    const float vl = __shfl_up_sync(0xffffffffu, V.v[3], 1), vr = __shfl_down_sync(0xffffffffu, V.v[0], 1);
    G.l = first ? G.v[0] : need_l ? w.gl : gl; G.r = last ? G.v[3] : need_r ? w.gr : gr;
    V.l = first ? V.v[0] : need_l ? w.vl : vl; V.r = last ? V.v[3] : need_r ? w.vr : vr;
}
__device__ __forceinline__ void dw_acc(const float (&w)[9], int dy, const SegRow& s, float (&o)[4]) {
    const float e[6] = {s.l, s.v[0], s.v[1], s.v[2], s.v[3], s.r};
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] += w[3 * dy] * e[j] + w[3 * dy + 1] * e[j + 1] + w[3 * dy + 2] * e[j + 2];
}

// one item = 4 columns x DW_BAND rows of one gate / value channel pair; consecutive threads = consecutive column quads (the
// lanes of a warp exchange the scalars at their quads' edges), then channel pairs (so that the CTAs in flight share the same
// rows of rs in L2), then bands, then images.  The operands of row r+2 are loaded while row r is computed.
template <bool ONE>
__global__ void __launch_bounds__(128) k_dwconv_gate(DwArgs a) {
    const int Q = a.W / 4, lane = (int)(threadIdx.x & 31u);
    const long total = (long)a.B * a.n_bands * a.Hd * Q, padded = (total + 31) & ~31L;       // whole warps iterate together
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < padded; i0 += (long)gridDim.x * blockDim.x) {
        const bool act = i0 < total;
        const long i = act ? i0 : total - 1;
        const int q = (int)(i % Q), k = (int)((i / Q) % a.Hd);
        const int band = (int)((i / ((long)Q * a.Hd)) % a.n_bands), b = (int)(i / ((long)Q * a.Hd * a.n_bands));
        const int r0 = band * DW_BAND, r1 = r0 + DW_BAND < a.H ? r0 + DW_BAND : a.H;
        const bool first = q == 0, last = q == Q - 1;
        const bool need_l = lane == 0 && !first, need_r = lane == 31 && !last;
        float wg[9], wv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) { wg[t] = a.w9[k * 9 + t]; wv[t] = a.w9[(a.Hd + k) * 9 + t]; }
        SegRow gU, gC, gD, vU, vC, vD;
        {
            const RawRow wU = dw_load<ONE>(a, b, k, r0 - 1, q, need_l, need_r), wC = dw_load<ONE>(a, b, k, r0, q, need_l, need_r);
            dw_finish<ONE>(wU, first, last, need_l, need_r, gU, vU);
            dw_finish<ONE>(wC, first, last, need_l, need_r, gC, vC);
        }
        RawRow nxt = dw_load<ONE>(a, b, k, r0 + 1, q, need_l, need_r);
        float* dst = a.u + (((long)b * a.Hd + k) * a.H + r0) * a.W + 4 * q;
        // every lane of a warp walks the same number of rows (its lanes may sit in different bands: the shuffles need them all)
        const int rend = r0 + (a.H < DW_BAND ? a.H : DW_BAND);
        for (int r = r0; r < rend; ++r, dst += a.W) {
            const RawRow nxt2 = dw_load<ONE>(a, b, k, r + 2, q, need_l, need_r);              // rows past H clamp: harmless re-reads
            dw_finish<ONE>(nxt, first, last, need_l, need_r, gD, vD);
            float G[4] = {0.f, 0.f, 0.f, 0.f}, V[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
            dw_acc(wg, 0, gU, G); dw_acc(wg, 1, gC, G); dw_acc(wg, 2, gD, G);
            dw_acc(wv, 0, vU, V); dw_acc(wv, 1, vC, V); dw_acc(wv, 2, vD, V);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = G[j] / (1.f + expf(-G[j])) * V[j];
            if (act && r < r1) st4(dst, o);
            gU = gC; gC = gD; vU = vC; vC = vD;
            nxt = nxt2;
        }
    }
}

static inline bool misaligned16(const void* p) { return ((uintptr_t)p & 15u) != 0; }

// x [B, C, H*W] -> rs [B, nsub, H*W] = 1 / sqrt(unbiased variance over each sub-net's C / nsub channels + eps).  HW % 4 == 0.
extern "C" int glrgtv_pixel_rstd(int B, int C, int nsub, long HW, float eps, const float* x, float* rs, void* stream) {
    if (B <= 0 || C <= 0 || nsub <= 0 || HW <= 0 || C % nsub || C / nsub < 2) return GLRGTV_ERR_SHAPE;
    if (HW % 4) return GLRGTV_ERR_UNSUPPORTED;
    if (!x || !rs || misaligned16(x) || misaligned16(rs)) return GLRGTV_ERR_POINTER;
    const long total = (long)B * nsub * (HW / 4), blocks = (total + 255) / 256;
    GLR_LAUNCH_FIBERS(k_pixel_rstd, dim3((unsigned)(blocks > 148 * 64 ? 148 * 64 : blocks)), 256, 0, stream, x, rs, B, nsub, C / nsub, HW, eps);
    return GLR_CHECK_LAUNCH();
}

// h [B, 2Hd, H, W] (un-normalised 1x1 output), rs [B, nsub, H, W], w9 [2Hd, 9] (the depthwise 3x3 weights), top / bot
// [B, 2Hd, W] already scaled neighbour rows or NULL (replicate) -> u [B, Hd, H, W].  W % 4 == 0, (2 Hd) % nsub == 0.
extern "C" int glrgtv_dwconv_gate(int B, int Hd, int nsub, int H, int W, const float* h, const float* rs, const float* w9,
                                  const float* top, const float* bot, float* u, void* stream) {
    if (B <= 0 || Hd <= 0 || nsub <= 0 || H <= 0 || W <= 0 || (2 * Hd) % nsub) return GLRGTV_ERR_SHAPE;
    if (W % 4) return GLRGTV_ERR_UNSUPPORTED;
    if (!h || !rs || !w9 || !u || misaligned16(h) || misaligned16(rs) || misaligned16(u) || (top && misaligned16(top)) || (bot && misaligned16(bot)))
        return GLRGTV_ERR_POINTER;
    DwArgs a;
    a.h = h; a.rs = rs; a.w9 = w9; a.top = top; a.bot = bot; a.u = u;
    a.B = B; a.Hd = Hd; a.nsub = nsub; a.H = H; a.W = W; a.n_bands = (H + DW_BAND - 1) / DW_BAND;
    const long total = (long)B * a.n_bands * Hd * (W / 4), blocks = (total + 127) / 128;
    if (blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    if (nsub == 1) GLR_LAUNCH_FIBERS(k_dwconv_gate<true>, dim3((unsigned)blocks), 128, 0, stream, a);
    else GLR_LAUNCH_FIBERS(k_dwconv_gate<false>, dim3((unsigned)blocks), 128, 0, stream, a);
    return GLR_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------------------------------------
// Backward of the same pieces (training; host_cnn.py's autograd function).  With s = rs * h (the scaled 1x1 output),
// m = dw3x3_replicate(s), u = sigmoid(g) g v:
//   k_dwgate_bwd_point :  recompute g | v, gM = dL/dm from gu (both halves), and the depthwise weight gradient
//                         gw9[ch][t] += sum_p gM[ch][p] s[ch][cl(p + o_t)]                    (h, rs, gu read once; gM written)
//   k_dwconv_bwd_input :  gs = adjoint of the replicate-padded convolution applied to gM (taps that fell into the padding fold
//                         back onto the border pixel), gh = rs * gs                             (gM, rs read; gh written)
//   k_pixel_norm_bwd   :  gx = s0 gout + gx1 - <gx1, x>_c rs^2 (x - mean_c x) / (c - 1): gx1 = W1'^T gh (the caller's GEMM) is the
//                         gradient through the 1x1; the gradient through rs needs no pass over the 2Hd hidden channels because
//                         sum_ch gs h = <W1'^T gs, x> = <gx1, x> / rs.
// ---------------------------------------------------------------------------------------------------------------------------
struct DwBwdArgs {
    const float *h, *rs, *w9, *gu;
    float *gM, *gw9;
    int B, Hd, nsub, H, W, n_bands;
};

// add `v` of every lane to *dst: one atomic per warp when all lanes share `key` (the usual case: a warp inside one plane row),
// one per lane otherwise
__device__ __forceinline__ void keyed_warp_atomic(float v, float* dst, bool uniform, int lane) {
    if (uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v != 0.f) atomicAdd(dst, v);
    } else if (v != 0.f) {
        atomicAdd(dst, v);
    }
}

template <bool ONE>
__global__ void __launch_bounds__(128) k_dwgate_bwd_point(DwBwdArgs a) {
    DwArgs f;
    f.h = a.h; f.rs = a.rs; f.w9 = a.w9; f.top = nullptr; f.bot = nullptr; f.u = nullptr;
    f.B = a.B; f.Hd = a.Hd; f.nsub = a.nsub; f.H = a.H; f.W = a.W; f.n_bands = a.n_bands;
    const int Q = a.W / 4, lane = (int)(threadIdx.x & 31u);
    const long total = (long)a.B * a.n_bands * a.Hd * Q, padded = (total + 31) & ~31L;
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < padded; i0 += (long)gridDim.x * blockDim.x) {
        const bool act = i0 < total;
        const long i = act ? i0 : total - 1;
        const int q = (int)(i % Q), k = (int)((i / Q) % a.Hd);
        const int band = (int)((i / ((long)Q * a.Hd)) % a.n_bands), b = (int)(i / ((long)Q * a.Hd * a.n_bands));
        const int r0 = band * DW_BAND, r1 = r0 + DW_BAND < a.H ? r0 + DW_BAND : a.H;
        const bool first = q == 0, last = q == Q - 1;
        const bool need_l = lane == 0 && !first, need_r = lane == 31 && !last;
        float wg[9], wv[9], ag[9], av[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) { wg[t] = a.w9[k * 9 + t]; wv[t] = a.w9[(a.Hd + k) * 9 + t]; ag[t] = av[t] = 0.f; }
        SegRow gU, gC, gD, vU, vC, vD;
        {
            const RawRow wU = dw_load<ONE>(f, b, k, r0 - 1, q, need_l, need_r), wC = dw_load<ONE>(f, b, k, r0, q, need_l, need_r);
            dw_finish<ONE>(wU, first, last, need_l, need_r, gU, vU);
            dw_finish<ONE>(wC, first, last, need_l, need_r, gC, vC);
        }
        const long pg = (((long)b * 2 * a.Hd + k) * a.H + r0) * a.W + 4 * q, pv = pg + (long)a.Hd * a.H * a.W;
        const float* gup = a.gu + (((long)b * a.Hd + k) * a.H + r0) * a.W + 4 * q;
        const int rend = r0 + (a.H < DW_BAND ? a.H : DW_BAND);
        for (int r = r0; r < rend; ++r) {
            const RawRow nxt = dw_load<ONE>(f, b, k, r + 1, q, need_l, need_r);
            dw_finish<ONE>(nxt, first, last, need_l, need_r, gD, vD);
            if (act && r < r1) {
                float G[4] = {0.f, 0.f, 0.f, 0.f}, V[4] = {0.f, 0.f, 0.f, 0.f}, gu[4], dG[4], dV[4];
                dw_acc(wg, 0, gU, G); dw_acc(wg, 1, gC, G); dw_acc(wg, 2, gD, G);
                dw_acc(wv, 0, vU, V); dw_acc(wv, 1, vC, V); dw_acc(wv, 2, vD, V);
                ld4(gup + (long)(r - r0) * a.W, gu);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float sg = 1.f / (1.f + expf(-G[j]));
                    dG[j] = gu[j] * V[j] * sg * (1.f + G[j] * (1.f - sg));       // d/dG of sigmoid(G) G
                    dV[j] = gu[j] * sg * G[j];
                }
                st4(a.gM + pg + (long)(r - r0) * a.W, dG);
                st4(a.gM + pv + (long)(r - r0) * a.W, dV);
                const SegRow* gs[3] = {&gU, &gC, &gD};
                const SegRow* vs[3] = {&vU, &vC, &vD};
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const float eg[6] = {gs[dy]->l, gs[dy]->v[0], gs[dy]->v[1], gs[dy]->v[2], gs[dy]->v[3], gs[dy]->r};
                    const float ev[6] = {vs[dy]->l, vs[dy]->v[0], vs[dy]->v[1], vs[dy]->v[2], vs[dy]->v[3], vs[dy]->r};
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { ag[3 * dy + dx] += dG[j] * eg[j + dx]; av[3 * dy + dx] += dV[j] * ev[j + dx]; }
                }
            }
            gU = gC; gC = gD; vU = vC; vC = vD;
        }
        // depthwise weight gradient: lanes of a warp usually share the channel pair
        const float k0 = __shfl_sync(0xffffffffu, (float)k, 0);
        float diff = (act && (float)k == k0) ? 0.f : 1.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
        const bool uniform = diff == 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            keyed_warp_atomic(act ? ag[t] : 0.f, a.gw9 + k * 9 + t, uniform, lane);
            keyed_warp_atomic(act ? av[t] : 0.f, a.gw9 + (a.Hd + k) * 9 + t, uniform, lane);
        }
    }
}

// one item = 4 columns x DW_BAND rows of ONE hidden channel: gs = conv^T(gM) with the replicate padding's border folding
__global__ void __launch_bounds__(128) k_dwconv_bwd_input(const float* __restrict__ gM, const float* __restrict__ rs, const float* __restrict__ w9,
                                                          float* __restrict__ gh, int B, int C2, int nsub, int H, int W, int n_bands) {
    const int Q = W / 4, lane = (int)(threadIdx.x & 31u);
    const long total = (long)B * n_bands * C2 * Q, padded = (total + 31) & ~31L;
    for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < padded; i0 += (long)gridDim.x * blockDim.x) {
        const bool act = i0 < total;
        const long i = act ? i0 : total - 1;
        const int q = (int)(i % Q), ch = (int)((i / Q) % C2);
        const int band = (int)((i / ((long)Q * C2)) % n_bands), b = (int)(i / ((long)Q * C2 * n_bands));
        const int r0 = band * DW_BAND, r1 = r0 + DW_BAND < H ? r0 + DW_BAND : H;
        const bool first = q == 0, last = q == Q - 1;
        const bool need_l = lane == 0 && !first, need_r = lane == 31 && !last;
        float w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = w9[ch * 9 + t];
        const float* plane = gM + ((long)b * C2 + ch) * H * W + 4 * q;
        const float* rsp = rs + ((long)b * nsub + ch / (C2 / nsub)) * H * W + 4 * q;
        // row of gM, ZERO outside the plane (rows and columns)
        auto load = [&](int row) {
            SegRow s;
            if (row < 0 || row >= H) {
                s.l = s.r = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) s.v[j] = 0.f;
            } else {
                const float* p = plane + (long)row * W;
                ld4(p, s.v);
                s.l = need_l ? p[-1] : 0.f;
                s.r = need_r ? p[4] : 0.f;
            }
            const float sl = __shfl_up_sync(0xffffffffu, s.v[3], 1), sr = __shfl_down_sync(0xffffffffu, s.v[0], 1);
            if (!first && !need_l) s.l = sl;
            if (!last && !need_r) s.r = sr;
            return s;
        };
        // contribution of source row `s` through kernel row `dy` (0..2 = offsets -1..+1): out[x] += w[dy][dx] s[x - (dx - 1)], plus
        // the taps of the border columns that were clamped onto themselves
        auto rowconv = [&](const SegRow& s, int dy, float (&o)[4]) {
            const float e[6] = {s.l, s.v[0], s.v[1], s.v[2], s.v[3], s.r};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] += w[3 * dy] * e[j + 2] + w[3 * dy + 1] * e[j + 1] + w[3 * dy + 2] * e[j];
            if (first) o[0] += w[3 * dy] * e[1];
            if (last) o[3] += w[3 * dy + 2] * e[4];
        };
        SegRow U = load(r0 - 1), C = load(r0);
        const int rend = r0 + (H < DW_BAND ? H : DW_BAND);
        for (int r = r0; r < rend; ++r) {
            const SegRow D = load(r + 1);
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            rowconv(D, 0, o);            // source row r+1 reached output row r through the tap dy = -1
            rowconv(C, 1, o);
            rowconv(U, 2, o);
            if (r == 0) rowconv(C, 0, o);            // row 0's upward taps were clamped onto row 0
            if (r == H - 1) rowconv(C, 2, o);
            if (act && r < r1) {
                float rv[4];
                ld4(rsp + (long)r * W, rv);
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] *= rv[j];
                st4(gh + ((long)b * C2 + ch) * H * W + (long)r * W + 4 * q, o);
            }
            U = C; C = D;
        }
    }
}

__global__ void __launch_bounds__(256) k_pixel_norm_bwd(const float* __restrict__ x, const float* __restrict__ rs, const float* __restrict__ gx1,
                                                        const float* __restrict__ gout, const float* __restrict__ s0p, float* __restrict__ gx,
                                                        int B, int nsub, int c, long HW) {
    const long Q = HW / 4, total = (long)B * nsub * Q;
    const float s0 = *s0p, invc = 1.f / (float)c, invc1 = 1.f / (float)(c - 1);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long q = i % Q, bs = i / Q;
        const long base = bs * c * HW + 4 * q;
        float sum[4] = {0.f, 0.f, 0.f, 0.f}, D[4] = {0.f, 0.f, 0.f, 0.f}, rv[4];
        for (int k = 0; k < c; ++k) {
            float xv[4], gv[4];
            ld4(x + base + (long)k * HW, xv);
            ld4(gx1 + base + (long)k * HW, gv);
#pragma unroll
            for (int j = 0; j < 4; ++j) { sum[j] += xv[j]; D[j] += xv[j] * gv[j]; }
        }
        ld4(rs + bs * HW + 4 * q, rv);
        float coef[4], mean[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { mean[j] = sum[j] * invc; coef[j] = D[j] * rv[j] * rv[j] * invc1; }
        for (int k = 0; k < c; ++k) {
            float xv[4], gv[4], go[4], o[4];
            ld4(x + base + (long)k * HW, xv);
            ld4(gx1 + base + (long)k * HW, gv);
            ld4(gout + base + (long)k * HW, go);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = s0 * go[j] + gv[j] - coef[j] * (xv[j] - mean[j]);
            st4(gx + base + (long)k * HW, o);
        }
    }
}

// gu [B,Hd,H,W] -> gM [B,2Hd,H,W] (scratch: dL/d(dw-conv output)), gh [B,2Hd,H,W] (dL/d(un-normalised 1x1 output)),
// gw9 [2Hd,9] += depthwise weight gradient (ACCUMULATES: the caller zeroes it).  Whole images only (no strip halos).
extern "C" int glrgtv_dwconv_gate_bwd(int B, int Hd, int nsub, int H, int W, const float* h, const float* rs, const float* w9,
                                      const float* gu, float* gM, float* gh, float* gw9, void* stream) {
    if (B <= 0 || Hd <= 0 || nsub <= 0 || H <= 0 || W <= 0 || (2 * Hd) % nsub) return GLRGTV_ERR_SHAPE;
    if (W % 4) return GLRGTV_ERR_UNSUPPORTED;
    if (!h || !rs || !w9 || !gu || !gM || !gh || !gw9 || misaligned16(h) || misaligned16(rs) || misaligned16(gu) || misaligned16(gM) || misaligned16(gh))
        return GLRGTV_ERR_POINTER;
    DwBwdArgs a;
    a.h = h; a.rs = rs; a.w9 = w9; a.gu = gu; a.gM = gM; a.gw9 = gw9;
    a.B = B; a.Hd = Hd; a.nsub = nsub; a.H = H; a.W = W; a.n_bands = (H + DW_BAND - 1) / DW_BAND;
    const long total = (long)B * a.n_bands * Hd * (W / 4), blocks = (total + 127) / 128;
    if (2 * blocks > 0x7fffffffL) return GLRGTV_ERR_SHAPE;
    if (nsub == 1) GLR_LAUNCH_FIBERS(k_dwgate_bwd_point<true>, dim3((unsigned)blocks), 128, 0, stream, a);
    else GLR_LAUNCH_FIBERS(k_dwgate_bwd_point<false>, dim3((unsigned)blocks), 128, 0, stream, a);
    GLR_LAUNCH_FIBERS(k_dwconv_bwd_input, dim3((unsigned)(2 * blocks)), 128, 0, stream, gM, rs, w9, gh, B, 2 * Hd, nsub, H, W, a.n_bands);
    return GLR_CHECK_LAUNCH();
}

// gx [B,C,HW] = s0 gout + gx1 - <gx1, x>_c rs^2 (x - mean_c x) / (c - 1) per sub-net of c = C / nsub channels; s0: device scalar.
extern "C" int glrgtv_pixel_norm_bwd(int B, int C, int nsub, long HW, const float* x, const float* rs, const float* gx1,
                                     const float* gout, const float* s0, float* gx, void* stream) {
    if (B <= 0 || C <= 0 || nsub <= 0 || HW <= 0 || C % nsub || C / nsub < 2) return GLRGTV_ERR_SHAPE;
    if (HW % 4) return GLRGTV_ERR_UNSUPPORTED;
    if (!x || !rs || !gx1 || !gout || !s0 || !gx || misaligned16(x) || misaligned16(rs) || misaligned16(gx1) || misaligned16(gout) || misaligned16(gx))
        return GLRGTV_ERR_POINTER;
    const long total = (long)B * nsub * (HW / 4), blocks = (total + 255) / 256;
    GLR_LAUNCH_FIBERS(k_pixel_norm_bwd, dim3((unsigned)(blocks > 148 * 64 ? 148 * 64 : blocks)), 256, 0, stream, x, rs, gx1, gout, s0, gx, B, nsub,
                      C / nsub, HW);
    return GLR_CHECK_LAUNCH();
}
