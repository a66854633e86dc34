// fw2.cu - switch and shape test of the second-generation forward stages (kernel: fw2.cuh; one fw2_<stage>.cu per stage)
#include "fw2.cuh"

// 0 (default): the round-1 forward kernels (block_stream_fwd.cu) - on a B200 they are as fast as these at the benchmark size
// and faster on 4K planes (profiles/r02_summary.md); 1: the forward entry points use the pair walkers where the shape allows
// and the caller supplied the scratch (glrgtv_block_saved.vc).  glrgtv_set_fwd_kernels(2) selects them.
int g_glr_fw2 = 0;
extern "C" int glrgtv_set_fwd_kernels(int generation) { g_glr_fw2 = generation == 2 ? 1 : 0; return GLRGTV_OK; }

// shapes the pair walkers take: W % 8 == 0 (16-byte rows at half resolution), H even; any width (column strips)
bool glr_fw2_eligible(const glrgtv_shape* s) {
    if ((s->W & 7) || (s->H & 1) || s->W < 16 || s->H < 4) return false;
    return f2_fits<FW_X2, false>(s) && f2_fits<FW_X2, true>(s) && f2_fits<FW_X3, false>(s) && f2_fits<FW_X3, true>(s);
}
