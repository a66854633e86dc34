// bw2_ba.cu - the BW_BA stage of the second-generation backward (kernel and launch templates: bw2.cuh)
#include "bw2.cuh"
template int glr_bw2_stage<BW_BA>(B2Args, const float*, const float*, float*, float*, float*, int, void*);
