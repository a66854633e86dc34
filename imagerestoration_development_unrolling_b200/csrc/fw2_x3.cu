// fw2_x3.cu - the FW_X3 stage of the second-generation forward (kernel and launch templates: fw2.cuh)
#include "fw2.cuh"
template int glr_fw2_stage<FW_X3>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
