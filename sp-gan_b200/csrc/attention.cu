// Row softmax (and its backward) for the generator's optional N x N attention block
// (--attn: Generation/modules.py:534-558, F.softmax(bmm(theta^T, phi), -1)).  One warp per row,
// the row is read twice (max, then exp-sum) and written once; rows are L2-resident (N <= a few thousand).
// Not a tuned path: --attn is a non-default flag ("compose from the same primitives").
#include "common.cuh"
#include <float.h>

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ x, int64_t R, int N, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const float* xr = x + r * N;
        float m = -FLT_MAX;
        for (int j = lane; j < N; j += 32) m = fmaxf(m, __ldg(xr + j));
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < N; j += 32) s += expf(__ldg(xr + j) - m);
        s = warp_sum(s);
        const float inv = 1.f / s;
        float* yr = y + r * N;
        for (int j = lane; j < N; j += 32) yr[j] = expf(__ldg(xr + j) - m) * inv;
    }
}

// dx = y * (g - sum_j g_j y_j)
__global__ void __launch_bounds__(256)
row_softmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t R, int N,
                       float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const float* gr = g + r * N;
        const float* yr = y + r * N;
        float s = 0.f;
        for (int j = lane; j < N; j += 32) s = fmaf(__ldg(gr + j), __ldg(yr + j), s);
        s = warp_sum(s);
        float* dr = dx + r * N;
        for (int j = lane; j < N; j += 32) dr[j] = __ldg(yr + j) * (__ldg(gr + j) - s);
    }
}

}  // namespace

extern "C" int spgan_row_softmax(const float* x, int64_t R, int N, float* y, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && y && R >= 0 && N >= 1);
    if (R == 0) return SPGAN_OK;
    row_softmax_kernel<<<ew_grid(R * 32, 256), 256, 0, as_stream(s)>>>(x, R, N, y);
    return spgan_launch_status();
}

extern "C" int spgan_row_softmax_bwd(const float* g, const float* y, int64_t R, int N, float* dx, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && y && dx && R >= 0 && N >= 1);
    if (R == 0) return SPGAN_OK;
    row_softmax_bwd_kernel<<<ew_grid(R * 32, 256), 256, 0, as_stream(s)>>>(g, y, R, N, dx);
    return spgan_launch_status();
}
