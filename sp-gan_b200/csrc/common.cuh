// Shared helpers for libspgan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spgan_b200.h"

#define SPGAN_CHECK_ARG(cond) \
    do {                      \
        if (!(cond)) return SPGAN_E_BADARG; \
    } while (0)

// Kernel launches report through the sticky-free peek so that a bad launch configuration
// surfaces as a return code instead of poisoning later calls.
static inline int spgan_launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return (int)e;
    }
    return SPGAN_OK;
}

static inline cudaStream_t as_stream(spgan_stream_t s) { return (cudaStream_t)s; }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// B200: 148 SMs.  Elementwise / reduction grids are sized in multiples of this.
constexpr int kNumSMs = 148;

static inline int ew_grid(int64_t work_items, int threads, int max_waves = 8) {
    int64_t blocks = ceil_div64(work_items, threads);
    int64_t cap = (int64_t)kNumSMs * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float lrelu_f(float x, float slope) { return x > 0.f ? x : x * slope; }
