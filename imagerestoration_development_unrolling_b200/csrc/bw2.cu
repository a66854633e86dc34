// bw2.cu - switches and shape test of the second-generation backward stages (kernel: bw2.cuh; one bw2_<stage>.cu per stage)
#include "bw2.cuh"

// 1 (default): glrgtv_block_bwd uses these kernels where the shape allows; 0: the round-1 kernels (a test / comparison switch)
int g_glr_bw2 = 1;
extern "C" int glrgtv_set_bwd_kernels(int generation) { g_glr_bw2 = generation == 1 ? 0 : 1; return GLRGTV_OK; }

// shapes the pair walkers take: W % 8 == 0 (16-byte rows at half resolution), H even, a CTA of <= B2_MAXT threads holds at least
// half of a graph's channels
bool glr_bw2_eligible(const glrgtv_shape* s) {
    if ((s->W & 7) || (s->H & 1) || s->W < 8 || s->H < 4) return false;
    return b2_fits<BW_X2A, false>(s) && b2_fits<BW_X2A, true>(s) && b2_fits<BW_X2B, false>(s) && b2_fits<BW_X2B, true>(s);
}
