// fw2_ba.cu - the FW_BA stage of the second-generation forward (kernel and launch templates: fw2.cuh)
#include "fw2.cuh"
template int glr_fw2_stage<FW_BA>(F2Args, const float*, const float*, const float*, float*, int, int, void*);
