// Pairwise Chamfer distance between two SETS of point clouds (BASELINE configs[4]):
//   cd[i, j] = mean_n min_m |a_i[n] - b_j[m]|^2 + mean_m min_n |a_i[n] - b_j[m]|^2
// which is what metrics/evaluation_metrics.py:89-125 (_pairwise_EMD_CD_: dl.mean(1) + dr.mean(1) of
// distChamfer, :37-49) and Common/GAN_metrics.py:658-684 (pairwise_CD) compute one reference-batch at a time.
// The arithmetic is the direct form (dx*dx + dy*dy + dz*dz in fp32) of the reference's own CUDA kernel
// (metrics/CD_EMD/cd/chamferdist/chamfer.cu:12-134), not the cancellation-prone expanded form of the torch
// fallback.
//
// One CTA per cloud pair: both clouds are staged in shared memory as structure-of-arrays (24 KB each at
// N = 2048), 256 threads as a 16 x 16 grid of 8 x 8 register tiles: every step evaluates 128 x 128 point
// pairs with 8 flops each and no memory traffic; row minima live in registers across the sweep over b,
// column minima are combined with one shuffle and a shared-memory atomicMin per 64 pairs.  HBM/L2 traffic is
// 48 KB per 4.2 M point pairs: the kernel is bound by fp32 issue, 8 instructions per point pair.
#include "common.cuh"
#include <float.h>

namespace {

constexpr int CH_THREADS = 256;
constexpr int TP = 8;                         // points per thread along each axis
constexpr int TILE = 16 * TP;                 // 128 points per CTA step along each axis
constexpr float FAR_A = 1.0e18f, FAR_B = -1.0e18f;   // padding: never a minimum, never overflows (d ~ 1e37)

__device__ __forceinline__ int pad_up(int n) { return (n + TILE - 1) / TILE * TILE; }

__global__ void __launch_bounds__(CH_THREADS)
pairwise_chamfer_kernel(const float* __restrict__ a, const float* __restrict__ b, int S, int R, int N, int M,
                        int64_t pair0, int64_t npairs, float* __restrict__ cd, float* __restrict__ dl,
                        float* __restrict__ dr) {
    extern __shared__ __align__(16) float sm[];
    const int Np = pad_up(N), Mp = pad_up(M);
    float* ax = sm;            float* ay = ax + Np;  float* az = ay + Np;
    float* bx = az + Np;       float* by = bx + Mp;  float* bz = by + Mp;
    int* colmin = reinterpret_cast<int*>(bz + Mp);
    __shared__ float red[CH_THREADS / 32], red2[CH_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    for (int64_t pair = pair0 + blockIdx.x; pair < pair0 + npairs; pair += gridDim.x) {
        const int i = (int)(pair / R), j = (int)(pair % R);
        const float* ai = a + (int64_t)i * N * 3;
        const float* bj = b + (int64_t)j * M * 3;
        __syncthreads();                                   // previous pair fully consumed
        for (int e = tid; e < Np; e += CH_THREADS) {
            const bool v = e < N;
            ax[e] = v ? __ldg(ai + 3 * e) : FAR_A; ay[e] = v ? __ldg(ai + 3 * e + 1) : FAR_A; az[e] = v ? __ldg(ai + 3 * e + 2) : FAR_A;
        }
        for (int e = tid; e < Mp; e += CH_THREADS) {
            const bool v = e < M;
            bx[e] = v ? __ldg(bj + 3 * e) : FAR_B; by[e] = v ? __ldg(bj + 3 * e + 1) : FAR_B; bz[e] = v ? __ldg(bj + 3 * e + 2) : FAR_B;
            colmin[e] = 0x7f800000;                        // +inf
        }
        __syncthreads();

        float rowsum = 0.f;
        for (int x0 = 0; x0 < Np; x0 += TILE) {
            float px[TP], py[TP], pz[TP], rmin[TP];
#pragma unroll
            for (int u = 0; u < TP; ++u) {
                const int e = x0 + ty * TP + u;
                px[u] = ax[e]; py[u] = ay[e]; pz[u] = az[e];
                rmin[u] = INFINITY;
            }
            for (int y0 = 0; y0 < Mp; y0 += TILE) {
                float qx[TP], qy[TP], qz[TP], cmin[TP];
                const int e0 = y0 + tx * TP;
#pragma unroll
                for (int v = 0; v < TP; v += 4) {
                    const float4 X = *reinterpret_cast<const float4*>(bx + e0 + v);
                    const float4 Y = *reinterpret_cast<const float4*>(by + e0 + v);
                    const float4 Z = *reinterpret_cast<const float4*>(bz + e0 + v);
                    qx[v] = X.x; qx[v + 1] = X.y; qx[v + 2] = X.z; qx[v + 3] = X.w;
                    qy[v] = Y.x; qy[v + 1] = Y.y; qy[v + 2] = Y.z; qy[v + 3] = Y.w;
                    qz[v] = Z.x; qz[v + 1] = Z.y; qz[v + 2] = Z.z; qz[v + 3] = Z.w;
                }
#pragma unroll
                for (int v = 0; v < TP; ++v) cmin[v] = INFINITY;
#pragma unroll
                for (int u = 0; u < TP; ++u)
#pragma unroll
                    for (int v = 0; v < TP; ++v) {
                        const float dx = px[u] - qx[v], dy = py[u] - qy[v], dz = pz[u] - qz[v];
                        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        rmin[u] = fminf(rmin[u], d);
                        cmin[v] = fminf(cmin[v], d);
                    }
                // column minima: the two ty groups of this warp share the same 128 b-points
#pragma unroll
                for (int v = 0; v < TP; ++v) {
                    const float o = __shfl_xor_sync(0xffffffffu, cmin[v], 16);
                    const float m = fminf(cmin[v], o);
                    if (lane < 16) atomicMin(colmin + e0 + v, __float_as_int(m));     // d >= 0: int order == float order
                }
            }
            // row minima: combine the 16 tx lanes that share this thread's 8 a-points
#pragma unroll
            for (int u = 0; u < TP; ++u) {
                float m = rmin[u];
                m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                if (tx == 0 && x0 + ty * TP + u < N) rowsum += m;
            }
        }
        __syncthreads();                                   // all atomicMin done
        float colsum = 0.f;
        for (int e = tid; e < M; e += CH_THREADS) colsum += __int_as_float(colmin[e]);
        // block sums (fixed order: deterministic)
        float rs = rowsum, cs = colsum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); cs += __shfl_xor_sync(0xffffffffu, cs, o); }
        if (lane == 0) { red[warp] = rs; red2[warp] = cs; }
        __syncthreads();
        if (tid == 0) {
            float R1 = 0.f, C1 = 0.f;
            for (int w = 0; w < CH_THREADS / 32; ++w) { R1 += red[w]; C1 += red2[w]; }
            const float l = R1 / (float)N, r = C1 / (float)M;
            const int64_t o = pair - pair0;
            if (cd) cd[o] = l + r;
            if (dl) dl[o] = l;
            if (dr) dr[o] = r;
        }
    }
}

}  // namespace

extern "C" int spgan_pairwise_chamfer(const float* a, const float* b, int S, int R, int N, int M, int64_t pair0,
                                      int64_t npairs, float* cd, float* dl, float* dr, spgan_stream_t s) {
    SPGAN_CHECK_ARG(a && b && (cd || dl || dr) && S >= 1 && R >= 1 && N >= 1 && M >= 1);
    SPGAN_CHECK_ARG(pair0 >= 0 && npairs >= 0 && pair0 + npairs <= (int64_t)S * R);
    if (npairs == 0) return SPGAN_OK;
    const int Np = (N + TILE - 1) / TILE * TILE, Mp = (M + TILE - 1) / TILE * TILE;
    const size_t smem = ((size_t)3 * Np + (size_t)4 * Mp) * sizeof(float);
    if (smem > 200 * 1024) return SPGAN_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(pairwise_chamfer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    int64_t grid = (int64_t)kNumSMs * per_sm * 4;          // a few waves: pairs are equal-sized work items
    if (grid > npairs) grid = npairs;
    pairwise_chamfer_kernel<<<(unsigned)grid, CH_THREADS, smem, as_stream(s)>>>(a, b, S, R, N, M, pair0, npairs, cd, dl, dr);
    return spgan_launch_status();
}
