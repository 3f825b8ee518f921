// trs_internal.h — what the translation units of libtrs_b200.so share besides the public header (not exported).
#pragma once
#include <cuda_runtime.h>

#include "../../include/trs_b200.h"

#define TRS_HIDDEN __attribute__((visibility("hidden")))

TRS_HIDDEN int trs_i_fail(int code, const char* fmt, ...);          // sets trs_last_error(), returns code
TRS_HIDDEN int trs_i_cuda_fail(cudaError_t e, const char* what);
TRS_HIDDEN void trs_i_count_launches(int k);                        // feeds trs_kernel_launches()
TRS_HIDDEN int trs_i_ctx_device(const trs_ctx* ctx);
TRS_HIDDEN int trs_i_ctx_sm_count(const trs_ctx* ctx);

// Every entry point runs on its context's device and puts the caller's current device back before it returns: PyTorch shares the
// process and derives ITS current device from cudaGetDevice().
struct TrsDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit TrsDeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = err == cudaSuccess; }
    }
    ~TrsDeviceGuard() { if (switched) cudaSetDevice(prev); }
    TrsDeviceGuard(const TrsDeviceGuard&) = delete;
    TrsDeviceGuard& operator=(const TrsDeviceGuard&) = delete;
};
