// fw2_x1.cu - the FW_X1 stage of the second-generation forward (kernel and launch templates: fw2.cuh)
#include "fw2.cuh"
template int glr_fw2_stage<FW_X1>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
