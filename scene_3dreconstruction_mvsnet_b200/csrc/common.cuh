// Shared helpers for the sm_100a kernels of libmvsnet_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvsnet_b200.h"

namespace mvs {

int set_error(int code, const char *fmt, ...);
void count_launches(int n);

#define MVS_REQUIRE(cond, ...)                                            \
    do {                                                                  \
        if (!(cond)) return ::mvs::set_error(MVS_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

#define MVS_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return ::mvs::set_error(MVS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                    __FILE__, __LINE__);                                                 \
    } while (0)

#define MVS_LAUNCH_CHECK(nlaunch)                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                              \
        if (e__ != cudaSuccess)                                                                            \
            return ::mvs::set_error(MVS_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                                    __FILE__, __LINE__);                                                   \
        ::mvs::count_launches(nlaunch);                                                                    \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- internal cross-file entry points (all asynchronous on `st`) ----
// rt [B*(V-1)][12] = rot(9) | trans(3) from proj [B,V,4,4]   (module.py:107-109)
int compose_homographies(const float *proj, float *rt, int B, int V, cudaStream_t st);
int compose_homography_pairs(const float *src_proj, const float *ref_proj, float *rt, int B, cudaStream_t st);

}  // namespace mvs
