// fw2.cu - switch and shape test of the second-generation forward stages (kernel: fw2.cuh; one fw2_<stage>.cu per stage)
#include "fw2.cuh"

// Which forward stage kernels run: 0 (default) = automatic, 1 = always the round-1 quad walkers (block_stream_fwd.cu), 2 = the
// pair walkers wherever the shape allows and the caller supplied the scratch (glrgtv_block_saved.vc).  The automatic rule is the
// measured one (B200, profiles/r02_configs.md, bench shapes): the pair walkers win every stage on planes of at most 64 columns
// and the BA / X2 stages at 128 columns; the quad walkers win on wider planes, 4K column strips included.
int g_glr_fw2 = 0;
extern "C" int glrgtv_set_fwd_kernels(int generation) {
    if (generation < 0 || generation > 2) return GLRGTV_ERR_SHAPE;
    g_glr_fw2 = generation;
    return GLRGTV_OK;
}
bool glr_fw2_wanted(int mode, const glrgtv_shape* s) {
    if (g_glr_fw2 == 1) return false;
    if (g_glr_fw2 == 2) return true;
    return s->W <= 64 || (s->W <= 128 && (mode == FW_BA || mode == FW_X2));
}

// shapes the pair walkers take: W % 8 == 0 (16-byte rows at half resolution), H even; any width (column strips)
bool glr_fw2_eligible(const glrgtv_shape* s) {
    if ((s->W & 7) || (s->H & 1) || s->W < 16 || s->H < 4) return false;
    return f2_fits<FW_X2, false>(s) && f2_fits<FW_X2, true>(s) && f2_fits<FW_X3, false>(s) && f2_fits<FW_X3, true>(s);
}
