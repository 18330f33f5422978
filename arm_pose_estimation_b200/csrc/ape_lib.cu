// Library-level entry points of libape_b200: ABI version, error text, device probe, blob size.
#include "ape_common.cuh"
#include "ape_lstm_pack.h"

namespace ape {
thread_local cudaError_t g_last_err = cudaSuccess;
}

extern "C" int ape_abi_version(void) { return 6; }

extern "C" const char* ape_last_cuda_error(void) { return cudaGetErrorString(ape::g_last_err); }

extern "C" int ape_device_info(int* sm_count, int* smem_optin_bytes, int* cc_major, int* cc_minor) {
    int dev = 0, v = 0;
    APE_CUDA_TRY(cudaGetDevice(&dev));
    APE_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = v;
    APE_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem_optin_bytes) *smem_optin_bytes = v;
    int major = 0, minor = 0;
    APE_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    APE_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return major == 10 ? APE_OK : APE_ERR_NO_SM100;
}

extern "C" int ape_lstm_blob_floats(int I, int H, int L, int O, int64_t* floats) {
    if (!floats || I < 1 || H < 1 || L < 1 || O < 1) return APE_ERR_BAD_ARG;
    *floats = ape_pack_total_floats(I, H, L, O);
    return APE_OK;
}
