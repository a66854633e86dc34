// api_misc.cu - version / device / error plumbing of the C ABI.
#include "common.cuh"

#ifdef GLRGTV_EMU
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local float emu_smem_storage[64 * 1024];
thread_local float* emu_smem = emu_smem_storage;
extern "C" {
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return "emulation build"; }
int glrgtv_check_device(void) { return GLRGTV_OK; }
}
#else
#include <string.h>
static thread_local char g_last_error[256] = "";

int glr_record_launch_error(void) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return GLRGTV_OK;
    strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
    return GLRGTV_ERR_CUDA;
}

extern "C" {
int glrgtv_abi_version(void) { return GLRGTV_ABI_VERSION; }
const char* glrgtv_last_cuda_error(void) { return g_last_error; }
int glrgtv_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GLRGTV_ERR_DEVICE;
    return major == 10 ? GLRGTV_OK : GLRGTV_ERR_DEVICE;
}
}
#endif
