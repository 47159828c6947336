// Shared host-side plumbing for libmonorec_b200.so: thread-local error text, launch counting, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "../../include/monorec_b200.h"

namespace mr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// cost_volume.cu: launches the fused kernel for batch elements [b_begin, b_begin + b_count) of a B-element problem
int launch_cost_volume(const float* keyframe, const float* const* frames, const float* proj, const float* depths,
                       float* out_cv, float* out_sfcv, int B, int F, int D, int H, int W, float alpha,
                       const float* chan_w, int b_begin, int b_count, int gather_only, cudaStream_t stream,
                       void* sf_nhwc = nullptr, int sf_nhwc_dtype = 0);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MR_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

#define MR_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            ::mr::set_error(__VA_ARGS__);      \
            return MR_EINVAL;                  \
        }                                      \
    } while (0)

#define MR_CUDA(call)                                            \
    do {                                                         \
        int _rc = ::mr::check_cuda((call), #call);               \
        if (_rc != MR_OK) return _rc;                            \
    } while (0)

// post-launch check (does not synchronise)
#define MR_LAUNCH_CHECK(name)                                    \
    do {                                                         \
        ::mr::count_launch();                                    \
        int _rc = ::mr::check_cuda(cudaGetLastError(), name);    \
        if (_rc != MR_OK) return _rc;                            \
    } while (0)

}  // namespace mr
