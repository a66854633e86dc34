// bw2_x3.cu - the BW_X3 stage of the second-generation backward (kernel and launch templates: bw2.cuh)
#include "bw2.cuh"
template int glr_bw2_stage<BW_X3>(B2Args, const float*, const float*, float*, float*, float*, int, void*);
