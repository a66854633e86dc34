// proj_gemm.cu - the two feature projections of MixtureGTVGLR (V1X0:556-612, 712, 725) as tensor-core GEMMs.
//
// patchs_features_extraction00 is a 1x1 convolution, ...01 a 2x2 stride-2 convolution followed by a 1x1: all three are
//     Y[b] (M x N) = W (M x K) . X[b] (K x N),        N = pixels, K = input channels (x4 after space-to-depth)
// and their input gradient is  gX[b] (K x N) = W^T . gY[b].  The 1e-4 parity bar rules out single-pass TF32 (1e-3), and
// cuBLAS answers fp32 requests with SIMT kernels (21 % of the training step, profiles/r01_summary.md).  These GEMMs use
// the tensor cores with 3xTF32 error compensation (big*big + big*small + small*big, fp32-level accuracy): CUTLASS
// GemmBatched with arch::OpMultiplyAddFastF32, templates from the CUTLASS header tree vendored in this image, instantiated
// here for sm_100a.  The weight gradient stays on cuBLAS (one reduction over all pixels, split-K).
//
// MEASURED on B200 (tools/proj_times.py): correct to 2e-6, but SLOWER than cuBLAS' fp32 SIMT GEMM (0.93 vs 0.63 ms for
// the scale-0 forward GEMM, 19 GFLOP): the legacy mma.sync path runs TF32 far below tcgen05 rates on sm_100a and three
// passes of it lose to the fp32 pipe.  The modules therefore keep cuBLAS by default (PROJ_TENSOR_CORES = False); the
// kernel that would pay is a tcgen05 kind::tf32 one with TMEM accumulators, fused in front of k_block_weights.
#include "common.cuh"

#ifndef GLRGTV_EMU
#include "cutlass/cutlass.h"
#include "cutlass/gemm/device/gemm_batched.h"

namespace {
using RowMajor = cutlass::layout::RowMajor;
using ColMajor = cutlass::layout::ColumnMajor;
template <class LA, class TB, class WARP>
using Gemm3x = cutlass::gemm::device::GemmBatched<
    float, LA, float, RowMajor, float, RowMajor, float, cutlass::arch::OpClassTensorOp, cutlass::arch::Sm80, TB, WARP,
    cutlass::gemm::GemmShape<16, 8, 8>, cutlass::epilogue::thread::LinearCombination<float, 4, float, float>,
    cutlass::gemm::threadblock::GemmBatchedIdentityThreadblockSwizzle, 3, 4, 4, cutlass::arch::OpMultiplyAddFastF32>;
using TB128 = cutlass::gemm::GemmShape<128, 128, 16>;
using W128 = cutlass::gemm::GemmShape<64, 64, 16>;
using TB64 = cutlass::gemm::GemmShape<64, 128, 16>;
using W64 = cutlass::gemm::GemmShape<32, 64, 16>;

template <class G>
int run(int m, int n, int k, const float* A, int lda, const float* B, int ldb, long sB, float* C, int ldc, long sC, int batch,
        cudaStream_t st) {
    G op;
    typename G::Arguments args({m, n, k}, {A, lda}, 0, {B, ldb}, sB, {C, ldc}, sC, {C, ldc}, sC, {1.f, 0.f}, batch);
    if (op.can_implement(args) != cutlass::Status::kSuccess) return GLRGTV_ERR_UNSUPPORTED;
    if (op.initialize(args, nullptr, st) != cutlass::Status::kSuccess) return GLRGTV_ERR_CUDA;
    ++g_glr_launches;
    return op(st) == cutlass::Status::kSuccess ? GLRGTV_OK : GLRGTV_ERR_CUDA;
}
}  // namespace

// transpose_w == 0:  Y[b] (M x N) = W (M x K)   . X[b] (K x N)      W [M,K], X [batch,K,N], Y [batch,M,N]
// transpose_w == 1:  Y[b] (K x N) = W^T (K x M) . X[b] (M x N)      W [M,K], X [batch,M,N], Y [batch,K,N]
// all row-major and contiguous; M, N, K multiples of 4 (128-bit tensor-core operand loads)
extern "C" int glrgtv_proj_gemm(int transpose_w, int batch, int M, int N, int K, const float* W, const float* X, float* Y,
                                void* stream) {
    if (batch <= 0 || M <= 0 || N <= 0 || K <= 0 || (M & 3) || (N & 3) || (K & 3)) return GLRGTV_ERR_SHAPE;
    if (!W || !X || !Y || (((uintptr_t)W | (uintptr_t)X | (uintptr_t)Y) & 15u)) return GLRGTV_ERR_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (!transpose_w) {
        rc = M <= 64 ? run<Gemm3x<RowMajor, TB64, W64>>(M, N, K, W, K, X, N, (long)K * N, Y, N, (long)M * N, batch, st)
                     : run<Gemm3x<RowMajor, TB128, W128>>(M, N, K, W, K, X, N, (long)K * N, Y, N, (long)M * N, batch, st);
    } else {
        // W [M,K] row-major read as W^T [K,M] column-major, leading dimension K
        rc = K <= 64 ? run<Gemm3x<ColMajor, TB64, W64>>(K, N, M, W, K, X, N, (long)M * N, Y, N, (long)K * N, batch, st)
                     : run<Gemm3x<ColMajor, TB128, W128>>(K, N, M, W, K, X, N, (long)M * N, Y, N, (long)K * N, batch, st);
    }
    return rc ? rc : GLR_CHECK_LAUNCH();
}
#endif
