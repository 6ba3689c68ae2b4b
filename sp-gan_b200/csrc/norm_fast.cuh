// Vectorised (float4) column reductions and per-column maps for row-major [R, C] tensors with
// C % 4 == 0 and 16-byte aligned pointers.  One thread owns one float4 column group for a chunk of
// rows, so per-column parameters (mean, rstd, gamma, ...) are loaded once and stay in registers.
#pragma once
#include "common.cuh"

namespace fastnorm {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
#define F4OP(name, expr)                                                              \
    __device__ __forceinline__ float4 name(float4 a, float4 b) {                       \
        return make_float4(expr(a.x, b.x), expr(a.y, b.y), expr(a.z, b.z), expr(a.w, b.w)); \
    }
#define ADD_(a, b) ((a) + (b))
#define SUB_(a, b) ((a) - (b))
#define MUL_(a, b) ((a) * (b))
F4OP(add4, ADD_)
F4OP(sub4, SUB_)
F4OP(mul4, MUL_)
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 mask4(float4 g, float4 y, float slope) {   // LeakyReLU backward by output sign
    return make_float4(y.x > 0.f ? g.x : g.x * slope, y.y > 0.f ? g.y : g.y * slope, y.z > 0.f ? g.z : g.z * slope,
                       y.w > 0.f ? g.w : g.w * slope);
}
__device__ __forceinline__ float4 lrelu4(float4 v, float s) {
    return make_float4(lrelu_f(v.x, s), lrelu_f(v.y, s), lrelu_f(v.z, s), lrelu_f(v.w, s));
}

struct Plan { int64_t nseg; int chunks; int64_t rows_per_chunk; };

// ---- reduction: partial[(blockIdx.y * NV + v) * C + c]
template <int NV, int TX, typename Op>
__global__ void __launch_bounds__(256)
colreduce4_kernel(Op op, int C4, int64_t seg_rows, int chunks, int64_t rows_per_chunk, float* __restrict__ partial) {
    constexpr int TY = 256 / TX;
    __shared__ float4 red[TY][NV][TX];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c4 = blockIdx.x * TX + tx;
    const int64_t seg = blockIdx.y / chunks;
    const int chunk = blockIdx.y % chunks;
    const int64_t rbeg = seg * seg_rows + (int64_t)chunk * rows_per_chunk;
    int64_t rend = rbeg + rows_per_chunk;
    if (rend > (seg + 1) * seg_rows) rend = (seg + 1) * seg_rows;
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = f4(0.f);
    if (c4 < C4) {
        const typename Op::State st = op.init(c4, seg);
        // Op::kStreams input tensors per row: keep ~8 float4 loads in flight per thread whatever the operator
        // (a one-stream reduction unrolled 4x ran at 4-5 TB/s next to 5.4-6 TB/s for the two-stream ones)
        constexpr int UNROLL = Op::kStreams >= 2 ? 4 : 8;
#pragma unroll UNROLL
        for (int64_t r = rbeg + ty; r < rend; r += TY) op.accum(st, r, c4, acc);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) red[ty][v][tx] = acc[v];
    __syncthreads();
    // threads with ty < NV finish value `ty` of column group tx
    if (ty < NV && c4 < C4) {
        float4 s = f4(0.f);
#pragma unroll
        for (int t = 0; t < TY; ++t) s = add4(s, red[t][ty][tx]);
        st4(partial + ((int64_t)blockIdx.y * NV + ty) * (C4 * 4) + c4 * 4, s);
    }
}

// ---- map
template <int TX, typename Op>
__global__ void __launch_bounds__(256)
colmap4_kernel(Op op, int C4, int64_t seg_rows, int chunks, int64_t rows_per_chunk) {
    constexpr int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c4 = blockIdx.x * TX + tx;
    if (c4 >= C4) return;
    const int64_t seg = blockIdx.y / chunks;
    const int chunk = blockIdx.y % chunks;
    const int64_t rbeg = seg * seg_rows + (int64_t)chunk * rows_per_chunk;
    int64_t rend = rbeg + rows_per_chunk;
    if (rend > (seg + 1) * seg_rows) rend = (seg + 1) * seg_rows;
    const typename Op::State st = op.init(c4, seg);
#pragma unroll 4
    for (int64_t r = rbeg + ty; r < rend; r += TY) op.apply(st, r, c4);
}

inline Plan make_plan(int64_t R, int C4, int TX, int64_t seg_rows, int target_blocks) {
    Plan p;
    p.nseg = R / seg_rows;
    const int64_t col_blocks = (C4 + TX - 1) / TX;
    int64_t want = (target_blocks + col_blocks * p.nseg - 1) / (col_blocks * p.nseg);
    const int TY = 256 / TX;
    int64_t maxc = (seg_rows + 4 * TY - 1) / (4 * TY);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    p.rows_per_chunk = ((seg_rows + want - 1) / want + TY - 1) / TY * TY;
    p.chunks = (int)((seg_rows + p.rows_per_chunk - 1) / p.rows_per_chunk);
    return p;
}

inline int pick_tx(int C4) { return C4 >= 32 ? 32 : (C4 >= 16 ? 16 : 8); }

template <typename Op>
int run_map(int64_t R, int C, int64_t seg_rows, cudaStream_t st, Op op) {
    const int C4 = C / 4, TX = pick_tx(C4);
    const Plan p = make_plan(R, C4, TX, seg_rows, 16 * kNumSMs);
    const int64_t gy = p.nseg * p.chunks;
    if (gy > 65535) return SPGAN_E_UNSUPPORTED;
    dim3 grid((C4 + TX - 1) / TX, (unsigned)gy);
    if (TX == 32) colmap4_kernel<32><<<grid, 256, 0, st>>>(op, C4, seg_rows, p.chunks, p.rows_per_chunk);
    else if (TX == 16) colmap4_kernel<16><<<grid, 256, 0, st>>>(op, C4, seg_rows, p.chunks, p.rows_per_chunk);
    else colmap4_kernel<8><<<grid, 256, 0, st>>>(op, C4, seg_rows, p.chunks, p.rows_per_chunk);
    return spgan_launch_status();
}

}  // namespace fastnorm
