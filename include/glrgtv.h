/*
 * glrgtv.h - C ABI of libglrgtv.so: hand-written sm_100a CUDA for the unrolled
 * graph-Laplacian-regulariser (GLR) / graph-total-variation (GTV) restoration blocks.
 *
 * The reference (tamthuc1995/ImageRestoration-Development-Unrolling) is pure PyTorch and has no FFI
 * of its own; every entry point below names the reference method it replaces.  V1X0 =
 * exploration/GGTV_GGLR_v1.0/deep_multiscale_GGLR_GGTV_v1x0.py (== LIB/model_GLR_GTV_deep_v13.py),
 * V7 = exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v7.py.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous float32 unless it says "host";
 *  - signals  are [B, G, F, H, W]   (== the reference's [B, C, H, W] with c = g*F + f);
 *    weights  are [B, G, E, H, W];  edge signals are [B, G, F, E, H, W];
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it, nothing synchronises,
 *    nothing allocates; outputs and workspaces belong to the caller;
 *  - "accumulated" outputs (parameter gradients) are ADDED to: the caller zeroes them;
 *  - return value: GLRGTV_OK or a negative glrgtv_status.  CUDA launch errors are reported through
 *    cudaGetLastError() -> GLRGTV_ERR_CUDA (glrgtv_last_cuda_error() gives the text).
 *  - there is no CPU fallback: host pointers are an error the driver will report at run time.
 */
#ifndef GLRGTV_H
#define GLRGTV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLRGTV_ABI_VERSION 2
#define GLRGTV_MAX_EDGES 48 /* 7x7 full window */

typedef enum glrgtv_status {
    GLRGTV_OK = 0,
    GLRGTV_ERR_SHAPE = -1,     /* non-positive / odd / inconsistent sizes        */
    GLRGTV_ERR_POINTER = -2,   /* NULL or misaligned pointer                     */
    GLRGTV_ERR_CUDA = -3,      /* launch failed, see glrgtv_last_cuda_error()    */
    GLRGTV_ERR_DEVICE = -4,    /* current device is not sm_100                   */
    GLRGTV_ERR_WORKSPACE = -5, /* workspace smaller than *_workspace_bytes()     */
    GLRGTV_ERR_UNSUPPORTED = -6
} glrgtv_status;

typedef enum glrgtv_pad { GLRGTV_PAD_CLAMP = 0, GLRGTV_PAD_REFLECT = 1 } glrgtv_pad;

/* signal geometry */
typedef struct glrgtv_shape {
    int32_t B, G, F, H, W;
} glrgtv_shape;

/* neighbour stencil: edge e connects p -> p + (dh[e], dw[e]); order as V1X0:42-49 / V7:284-298 */
typedef struct glrgtv_window {
    int32_t n_edges;
    int32_t dh[GLRGTV_MAX_EDGES];
    int32_t dw[GLRGTV_MAX_EDGES];
} glrgtv_window;

/* the four stats_kernel_p* parameters of one GLRFast/GTVFast (V1X0:66-118).  n = C (V1X0, per channel)
 * or 1 (V7:311-363, scalars shared by all channels).  pad = CLAMP for V1X0:186, REFLECT for V7:458 */
typedef struct glrgtv_stats {
    const float* p01;
    const float* p02a;
    const float* p02b;
    const float* p03;
    int32_t n;
    int32_t pad;
} glrgtv_stats;

int glrgtv_abi_version(void);
const char* glrgtv_last_cuda_error(void);
/* 0 when the current CUDA device can run this library (compute capability 10.x) */
int glrgtv_check_device(void);

/* kernels launched by this library since it was loaded (bench.py's gpu_launches) */
unsigned long long glrgtv_launch_count(void);

/* Optional per-kernel timing of the fused block entry points: one CUDA-event pair is recorded on the
 * launching stream around every kernel while enabled.  glrgtv_profile_read() waits for the events and
 * returns, per slot (glrgtv_prof_slot), the summed duration in ms and the number of launches. */
typedef enum glrgtv_prof_slot {
    GLRGTV_SLOT_FWD_WEIGHTS = 0,
    GLRGTV_SLOT_FWD_BA = 1, GLRGTV_SLOT_FWD_X1 = 2, GLRGTV_SLOT_FWD_X2 = 3, GLRGTV_SLOT_FWD_X3 = 4,
    GLRGTV_SLOT_BWD_X3 = 5, GLRGTV_SLOT_BWD_X2 = 6, GLRGTV_SLOT_BWD_X1 = 7, GLRGTV_SLOT_BWD_BA = 8,
    GLRGTV_SLOT_BWD_WEIGHTS = 9,
    /* the feature projections (csrc/proj_tc.cu + the space-to-depth copy): forward, input gradient, weight gradient */
    GLRGTV_SLOT_PROJ_FWD = 10, GLRGTV_SLOT_PROJ_DGRAD = 11, GLRGTV_SLOT_PROJ_WGRAD = 12,
    /* edge-weight gradient pass of the ROUND-1 backward (block_gw.cu; the default backward folds it into the stages) */
    GLRGTV_SLOT_GW = 13,
    GLRGTV_SLOT_COUNT = 14
} glrgtv_prof_slot;
int glrgtv_profile_enable(int on);
int glrgtv_profile_read(float* ms, int* count, int n_slots);

/* ------------------------------------------------------------------------------------------------
 * Per-operator entry points (public methods of GLRFast / GTVFast)
 * ---------------------------------------------------------------------------------------------- */

/* GLRFast/GTVFast.extract_edge_weights (V1X0:160-175, 391-407): feat [B,G,F,H,W], multiM [G,F] -> w [B,G,E,H,W] */
int glrgtv_edge_weights_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* feat,
                            const float* multiM, float* w, void* stream);
/* VJP: gw [B,G,E,H,W] -> gfeat [B,G,F,H,W] (written), gmultiM [G,F] (accumulated).
 * scratch: B*G*E*H*W floats (the softmax-VJP of gw). */
int glrgtv_edge_weights_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* feat,
                            const float* multiM, const float* w, const float* gw, float* gfeat,
                            float* gmultiM, float* scratch, void* stream);

/* normalize_and_transform_features (V1X0:146-157): out = multiM * feat / max(|feat|_F, 1e-12).
 * bwd: gfeat written, gmultiM accumulated. */
int glrgtv_normalize_fwd(const glrgtv_shape* s, const float* feat, const float* multiM, float* out, void* stream);
int glrgtv_normalize_bwd(const glrgtv_shape* s, const float* feat, const float* multiM, const float* g,
                         float* gfeat, float* gmultiM, void* stream);
/* get_neighbors_pixels (V1X0:128-144): out [B,G,F,E,H,W] = x[cl(p+d_e)]; bwd scatters back. */
int glrgtv_gather_neighbors_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, float* out,
                                void* stream);
int glrgtv_gather_neighbors_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* g, float* gx,
                                void* stream);

/* stats_conv (V1X0:177-195): out = S x.   bwd: gx written; gstats[4*n] accumulated, order p01,p02a,p02b,p03 */
int glrgtv_stats_conv_fwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* x, float* out, void* stream);
int glrgtv_stats_conv_bwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* x, const float* g,
                          float* gx, float* gstats, void* stream);
/* stats_conv_transpose (V1X0:197-215) */
int glrgtv_stats_conv_t_fwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* y, float* out, void* stream);
int glrgtv_stats_conv_t_bwd(const glrgtv_shape* s, const glrgtv_stats* st, const float* y, const float* g,
                            float* gy, float* gstats, void* stream);

/* GLRFast.op_L_norm (V1X0:218-228): out = x - sum_e w_e x[cl(p+d_e)].  bwd: gx, gw written. */
int glrgtv_op_L_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, const float* w,
                    float* out, void* stream);
int glrgtv_op_L_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* x, const float* w,
                    const float* g, float* gx, float* gw, void* stream);

/* GTVFast.op_C after its stats_conv (V1X0:459-467): z[b,g,f,e] = w_e (sx - sx[cl(p+d_e)]) */
int glrgtv_op_C_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* sx, const float* w,
                    float* z, void* stream);
int glrgtv_op_C_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* sx, const float* w,
                    const float* gz, float* gsx, float* gw, void* stream);
/* GTVFast.op_C_transpose before its stats_conv_transpose (V1X0:471-513) */
int glrgtv_op_Ct_fwd(const glrgtv_shape* s, const glrgtv_window* win, const float* z, const float* w,
                     float* o, void* stream);
int glrgtv_op_Ct_bwd(const glrgtv_shape* s, const glrgtv_window* win, const float* z, const float* w,
                     const float* go, float* gz, float* gw, void* stream);

/* MixtureGTVGLR.soft_threshold (V1X0:684-704): t [B,G,F,E,H,W], thr [G] (already exp'd) */
int glrgtv_soft_threshold_fwd(const glrgtv_shape* s, int n_edges, const float* t, const float* thr,
                              float* out, void* stream);
/* gt written, gthr [G] accumulated */
int glrgtv_soft_threshold_bwd(const glrgtv_shape* s, int n_edges, const float* t, const float* thr,
                              const float* g, float* gt, float* gthr, void* stream);

/* Mixture weighting of the older family (V7:847-858, 1011-1014): out[b,c] = sum_g x[b,g,c] * score[b,g];
 * x [B,G,F,H,W], score [B,G,H,W] (already soft-maxed over g), out [B,F,H,W].  bwd writes gx and gscore. */
int glrgtv_mixture_fwd(const glrgtv_shape* s, const float* x, const float* score, float* out, void* stream);
int glrgtv_mixture_bwd(const glrgtv_shape* s, const float* x, const float* score, const float* gout, float* gx,
                       float* gscore, void* stream);

/* 2x2 mean pooling P (V1X0:613, 662-665) and its transpose (V1X0:676-679); each is the other's VJP.
 * `s` is always the FINE geometry (H, W even); coarse tensors are [B,G,F,H/2,W/2]. */
int glrgtv_pool2_fwd(const glrgtv_shape* s, const float* fine, float* coarse, void* stream);
int glrgtv_unpool2_fwd(const glrgtv_shape* s, const float* coarse, float* fine, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused block: LocalLowpassFilteringBlock / MixtureGTVGLR (V1X0:707-811, 985-988), 3x3 cross window
 * ---------------------------------------------------------------------------------------------- */

/* parameters of one GLRFast / GTVFast inside the block */
typedef struct glrgtv_opparams {
    glrgtv_stats stats; /* n = C, pad = CLAMP */
    const float* multiM; /* [G,F] */
} glrgtv_opparams;

/* all parameters of MixtureGTVGLR that the stencil kernels read (the 1x1 / 2x2 projections are plain
 * GEMMs done by the caller; their outputs arrive as feat0 / feat1) */
typedef struct glrgtv_block_params {
    glrgtv_opparams gtv0, glr0, gtv1, glr1; /* GTVmodule00, GLRmodule00, GTVmodule01, GLRmodule01 */
    const float* alpha;  /* alphaCGD [3,G]          V1X0:546 */
    const float* beta;   /* betaCGD  [3,G]          V1X0:551 */
    const float* mu0;    /* muys00  [G] (log)       V1X0:582 */
    const float* ro0;    /* ro00    [G] (log)       V1X0:568 */
    const float* gamma0; /* gamma00 [G] (log)       V1X0:572 */
    const float* mu1;    /* muys01  [G] (log) */
    const float* ro1;    /* ro01    [G] (log) */
    const float* gamma1; /* gamma01 [G] (log) */
    const float* skip;   /* LocalLowpassFilteringBlock.skip_weight [2]; NULL => out = filter(x) */
} glrgtv_block_params;

/* gradients of the above, same layout, all ACCUMULATED (caller zeroes).  stats pointers are written
 * as mutable floats; n/pad are ignored. */
typedef struct glrgtv_block_grads {
    float *gtv0_stats, *glr0_stats, *gtv1_stats, *glr1_stats; /* each [4*C]: p01,p02a,p02b,p03 */
    float *gtv0_M, *glr0_M, *gtv1_M, *glr1_M;                 /* each [G,F] */
    float *alpha, *beta;                                      /* [3,G] */
    float *mu0, *ro0, *gamma0, *mu1, *ro1, *gamma1;           /* [G]   */
    float* skip;                                              /* [2] or NULL */
} glrgtv_block_grads;

/* activations kept between forward and backward (all written by glrgtv_block_fwd, caller-owned):
 *   w  : 4 weight sets  wT0,wL0 [B,G,4,H,W]  wT1,wL1 [B,G,4,H/2,W/2]
 *   bA, x1, bB, r1, x2 : [B,G,F,H,W] each                                   
 *   cT : symmetric GTV coefficients of wT (cR = wR^2 + wL[.,w+1]^2, cD = wD^2 + wU[h+1,.]^2):
 *        cT0 [B,G,2,H,W], cT1 [B,G,2,H/2,W/2]; read by the register-streaming stage kernels */
typedef struct glrgtv_block_saved {
    float *wT0, *wL0, *wT1, *wL1;
    float *bA, *x1, *bB, *r1, *x2;
    float *cT0, *cT1;
    /* scratch of the second-generation forward stage kernels (csrc/fw2.cuh): the half-resolution launches' results,
     * [2][B,G,F,H/2,W/2]; NULL => the round-1 forward kernels run */
    float* vc;
} glrgtv_block_saved;

/* Which kernels the fused block entry points use: 0 = automatic (register-streaming kernels when W % 8 == 0, planes
 * wider than 256 in column strips; shared-memory plane kernels otherwise), 1 = plane kernels only, 2 = streaming kernels only
 * (GLRGTV_ERR_UNSUPPORTED for other shapes).  A debugging / test switch; both paths compute the same function. */
int glrgtv_set_block_path(int mode);
/* Loader of the streaming kernels: 0 = automatic (per stage, as measured), 1 = per-thread cp.async rings,
 * 2 = one producer warp issuing TMA bulk row copies (cp.async.bulk + mbarrier).  Same results; a tuning switch. */
int glrgtv_set_stream_loader(int mode);
/* Backward stage kernels: 2 (default) = the pair walkers of csrc/bw2.cu (packed fp32 arithmetic, edge-weight gradients folded
 * into the adjoint walk; W % 4 == 0, all channels of a graph in one CTA), 1 = the round-1 walkers + separate gradient pass.
 * Same results; a test / comparison switch. */
int glrgtv_set_bwd_kernels(int generation);
/* Forward stage kernels: 0 (default) = automatic per stage and plane width, 1 = always the round-1 quad walkers of
 * csrc/block_stream_fwd.cu, 2 = the pair walkers of csrc/fw2.cuh wherever the shape allows (W % 8 == 0, any width in column
 * strips; they need the scratch glrgtv_block_saved.vc).  Same results; the automatic rule follows the B200 measurements in
 * profiles/r02_configs.md (pair walkers on planes of <= 64 columns and for the BA / X2 stages at 128 columns). */
int glrgtv_set_fwd_kernels(int generation);
/* Edge-weight construction and its VJP inside the block entry points: 0 (default) = the row walkers of csrc/weights_walk.cu
 * where the shape is theirs (W % 4 == 0, F = 6 or 12), 1 = always the round-1 tile kernels (kept as the second implementation
 * the walkers are tested against). */
int glrgtv_set_weights_kernels(int generation);
/* walker launches since the library was loaded (diagnostic: lets a test assert which kernels ran) */
unsigned long long glrgtv_weights_walk_launch_count(void);
/* streaming-path kernels launched since the library was loaded (diagnostic: lets a test assert which path ran) */
unsigned long long glrgtv_stream_launch_count(void);

/* x [B,C,H,W]; feat0 = patchs_features_extraction00(x) [B,2C,H,W]; feat1 = ..01(x) [B,2C,H/2,W/2]
 * (first C channels feed GTV, last C feed GLR, V1X0:714, 726);  out [B,C,H,W]. */
int glrgtv_block_fwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                     const float* feat0, const float* feat1, float* out, const glrgtv_block_saved* saved,
                     void* stream);

/* The same forward, one piece at a time, for callers that interleave the pieces with their own work - spatially sharded
 * inference exchanges halo rows between stages (shard.py).  stage 0: edge weights and GTV coefficients of the whole local
 * plane (needs feat0 / feat1); stages 1..4: BA (x -> saved.bA), X1 (bA -> x1), X2 (x1, x -> x2, bB, r1), X3 (x2, bB, r1,
 * x -> out) on the rows [row0, row1) only (even bounds).  A stage reads its stencil input (x | bA | x1 | x2) and the weights
 * up to 8 rows outside [row0, row1): those rows must hold valid data, or lie outside the plane (then the reference's border
 * rules apply).  Streaming kernels only: W % 8 == 0, else GLRGTV_ERR_UNSUPPORTED. */
int glrgtv_block_fwd_stage(int stage, const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                           const float* feat0, const float* feat1, float* out, const glrgtv_block_saved* saved,
                           int row0, int row1, void* stream);

size_t glrgtv_block_bwd_workspace_bytes(const glrgtv_shape* s);
/* gout [B,C,H,W] -> gx [B,C,H,W] (the direct path, WITHOUT the feature-path term),
 * gfeat0 [B,2C,H,W], gfeat1 [B,2C,H/2,W/2] (for the caller's GEMM backward), and parameter grads. */
int glrgtv_block_bwd(const glrgtv_shape* s, const glrgtv_block_params* p, const float* x,
                     const float* feat0, const float* feat1, const glrgtv_block_saved* saved,
                     const float* gout, float* gx, float* gfeat0, float* gfeat1,
                     const glrgtv_block_grads* grads, void* workspace, size_t workspace_bytes, void* stream);

/* Pipeline of the projection GEMM kernels: 0 (default) = tiles are split in place inside their pipeline stage, 1 = raw tiles
 * land in a deep ring of their own and are split into a shallow ring of operand slots (more HBM bytes in flight per SM;
 * measured no faster on a B200 - the kernel was bound by its epilogue, profiles/r02_proj_stalls.md - kept as a tested
 * alternative).  Same arithmetic, same results. */
int glrgtv_set_proj_pipeline(int which);
/* ------------------------------------------------------------------------------------------------
 * Feature projections patchs_features_extraction00 / 01 (V1X0:556-612, 712, 725; nn.Conv2d 1x1 and, after a
 * space-to-depth, 2x2 stride 2) on the tcgen05 tensor cores (csrc/proj_tc.cu): kind::tf32 MMAs with TMEM accumulators,
 * TMA operand loads, and the three-pass split a_hi b_hi + a_lo b_hi + a_hi b_lo (fp32-level accuracy, ~1e-6).
 * All operands row-major, contiguous, 16-byte aligned; M, N (pixels), K % 4 == 0 (tiles are zero-padded by the TMA unit).
 *   transpose_w == 0:  Y[b] (M x N) = W (M x K)   . X[b] (K x N)      forward
 *   transpose_w == 1:  Y[b] (K x N) = W^T (K x M) . X[b] (M x N)      input gradient
 * workspace: glrgtv_proj_gemm_workspace_bytes(M, K) bytes (the weights split into TF32 hi | lo parts and laid out as the
 * shared-memory tile images the kernel bulk-copies, one per accumulator chunk and 32-wide k-stage).
 * ---------------------------------------------------------------------------------------------- */
size_t glrgtv_proj_gemm_workspace_bytes(int M, int K);
int glrgtv_proj_gemm(int transpose_w, int batch, int M, int N, int K, const float* W, const float* X, float* Y,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Weight gradient of the same projections: gW [M,K] += sum_b gY[b] (M x N) . X[b]^T (N x K), same tensor-core scheme,
 * the reduction over batch and pixels split across the grid and reduced with red.global.add (ACCUMULATES: the caller
 * zeroes gW).  M, N, K % 4 == 0; GLRGTV_ERR_UNSUPPORTED otherwise. */
int glrgtv_proj_wgrad(int batch, int M, int N, int K, const float* gY, const float* X, float* gW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host CNN, inference forward: the memory-bound pieces of LocalNonLinearBlock (V1X0:911-964).
 * The caller does the two 1x1 convolutions as GEMMs (norm weight folded into the first, skip weights into the second).
 *
 * glrgtv_pixel_rstd: CustomLayerNorm's per-pixel scale (V1X0:918-925).  x [B,C,HW] -> rs [B,nsub,HW],
 *   rs = 1 / sqrt(var + eps), var = unbiased variance over the C/nsub channels of each sub-net.  HW % 4 == 0.
 * glrgtv_dwconv_gate: LocalGatedLinearBlock's depthwise 3x3 (replicate padded) and gate (V1X0:938-947) on the
 *   un-normalised 1x1 output h [B,2Hd,H,W]:  m = dw3x3(rs * h);  u [B,Hd,H,W] = sigmoid(g) * g * v with g = m[:, :Hd],
 *   v = m[:, Hd:].  w9 [2Hd,9] = the convolution weight [2Hd,1,3,3].  top / bot [B,2Hd,W]: the ALREADY SCALED rows
 *   above / below a row strip of a spatially sharded image, or NULL at the true image border (replicate).  W % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
int glrgtv_pixel_rstd(int B, int C, int nsub, long HW, float eps, const float* x, float* rs, void* stream);
int glrgtv_dwconv_gate(int B, int Hd, int nsub, int H, int W, const float* h, const float* rs, const float* w9,
                       const float* top, const float* bot, float* u, void* stream);

/* Backward of the two pieces above (training; host_cnn.py's autograd function), whole images only.
 * glrgtv_dwconv_gate_bwd: gu [B,Hd,H,W] = dL/du -> gh [B,2Hd,H,W] = dL/dh (h as given to the forward), gw9 [2Hd,9] +=
 *   depthwise weight gradient (ACCUMULATES: the caller zeroes it); gM [B,2Hd,H,W] is scratch (dL/d conv output).
 * glrgtv_pixel_norm_bwd: the input gradient of the block: gx [B,C,HW] = s0 gout + gx1 - <gx1,x>_c rs^2 (x - mean_c x)/(c-1),
 *   where gx1 = W1'^T gh is the caller's GEMM (gradient through the 1x1) and the last term is the gradient through rs;
 *   s0 points at the block's first skip weight on the device. */
int glrgtv_dwconv_gate_bwd(int B, int Hd, int nsub, int H, int W, const float* h, const float* rs, const float* w9,
                           const float* gu, float* gM, float* gh, float* gw9, void* stream);
int glrgtv_pixel_norm_bwd(int B, int C, int nsub, long HW, const float* x, const float* rs, const float* gx1,
                          const float* gout, const float* s0, float* gx, void* stream);

/* Space-to-depth in front of the 2x2 stride-2 projection (patchs_features_extraction01[0], V1X0:593-603) and its inverse:
 * inverse == 0: x [planes,H,W] -> y [planes*4,H/2,W/2], y[p*4 + dy*2 + dx, h, w] = x[p, 2h+dy, 2w+dx] (pixel_unshuffle order);
 * inverse == 1: the other way round (x is the deep tensor).  `planes` = B*C; W % 8 == 0, H even, 16-byte aligned. */
int glrgtv_space_to_depth(int inverse, long planes, int H, int W, const float* x, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLRGTV_H */
