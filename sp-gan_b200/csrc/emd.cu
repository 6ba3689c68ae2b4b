// Approximate Earth Mover's Distance (synchronous auction) for the evaluation path, SURVEY 8f-2.
// Replaces metrics/emd/emd_cuda.cu:95-282 (7 kernel launches per auction iteration, all state in global memory,
// called with eps = 0.005, 300 iterations at Common/GAN_metrics.py:375-379, 406-407) by ONE launch per batch of
// cloud pairs: a CTA owns a pair, keeps the target cloud, the prices and every assignment array in shared memory and
// runs all iterations itself (block barriers instead of kernel boundaries; stops as soon as nothing is unassigned).
// Arithmetic and tie rules are those of oracle/emd_recipe.c, so dist and assignment are bit-identical to the oracle:
//   value(j,k) = float(3.0 - double(sqrt(fma(dz,dz,fma(dy,dy,dx*dx)))) - double(price[k])), best = largest value
//   (lowest k on ties), better = second largest, bid increment = (best - better) + eps; among equal (within 1e-6)
//   top bids for one target the highest point index wins.
// tests/test_gpu_emd.py: dist and assignment bit-identical to the oracle on B200 (converged, unconverged, n not a
// multiple of 1024).
#include "common.cuh"

namespace {

constexpr int EMD_THREADS = 1024;
constexpr float NEG_BIG = -1e9f;

struct Top2 { float best, better; int i; };

__device__ __forceinline__ Top2 combine(Top2 a, Top2 b) {
    if (b.best > a.best || (b.best == a.best && b.i >= 0 && (a.i < 0 || b.i < a.i))) {
        a.better = fmaxf(a.best, b.better);
        a.best = b.best;
        a.i = b.i;
    } else {
        a.better = fmaxf(a.better, b.best);
    }
    return a;
}

__device__ __forceinline__ float bid_value(float x1, float y1, float z1, float x2, float y2, float z2, float price) {
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    const float sq = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));      // the reference binary's contraction order
    return __double2float_rn(3.0 - (double)__fsqrt_rn(sq) - (double)price);
}

__global__ void __launch_bounds__(EMD_THREADS, 1)
emd_auction_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n, float eps, int iters,
                   float* __restrict__ dist, int32_t* __restrict__ assignment) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* x2 = reinterpret_cast<float*>(smem_raw);
    float* y2 = x2 + n;
    float* z2 = y2 + n;
    float* price = z2 + n;
    float* bid_inc = price + n;
    int* max_inc = reinterpret_cast<int*>(bid_inc + n);       // float bits: every stored bid increment is > 0
    int* ass = max_inc + n;
    int* ass_inv = ass + n;
    int* bid = ass_inv + n;
    int* max_idx = bid + n;
    int* unass = max_idx + n;
    float* red_best = reinterpret_cast<float*>(unass + n);    // one Top2 per warp
    float* red_better = red_best + 32;
    int* red_i = reinterpret_cast<int*>(red_better + 32);
    __shared__ int s_nu;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p1 = xyz1 + (int64_t)blockIdx.x * n * 3;
    const float* p2 = xyz2 + (int64_t)blockIdx.x * n * 3;
    for (int k = tid; k < n; k += EMD_THREADS) {
        x2[k] = __ldg(p2 + k * 3 + 0); y2[k] = __ldg(p2 + k * 3 + 1); z2[k] = __ldg(p2 + k * 3 + 2);
        price[k] = 0.f; bid_inc[k] = 0.f; max_inc[k] = 0;
        ass[k] = -1; ass_inv[k] = -1; bid[k] = 0;
    }
    __syncthreads();

    for (int it = 0; it < iters; ++it) {
        const bool last = (it == iters - 1);
        if (tid == 0) s_nu = 0;
        for (int k = tid; k < n; k += EMD_THREADS) max_idx[k] = -1;
        __syncthreads();
        for (int j = tid; j < n; j += EMD_THREADS)
            if (ass[j] == -1) unass[atomicAdd(&s_nu, 1)] = j;           // order of the list does not affect results
        __syncthreads();
        const int nu = s_nu;
        if (nu == 0) break;                                             // converged: later iterations are no-ops

        // ---- Bid: tpp threads per unassigned point (power of two), targets interleaved over the group
        if (nu >= EMD_THREADS) {
            for (int u = tid; u < nu; u += EMD_THREADS) {
                const int j = unass[u];
                const float x1 = __ldg(p1 + j * 3 + 0), y1 = __ldg(p1 + j * 3 + 1), z1 = __ldg(p1 + j * 3 + 2);
                Top2 t{NEG_BIG, NEG_BIG, -1};
                for (int k = 0; k < n; ++k) {
                    const float d = bid_value(x1, y1, z1, x2[k], y2[k], z2[k], price[k]);
                    if (d > t.best) { t.better = t.best; t.best = d; t.i = k; }
                    else if (d > t.better) t.better = d;
                }
                const float inc = __fadd_rn(__fsub_rn(t.best, t.better), eps);
                bid[j] = t.i;
                bid_inc[j] = inc;
                atomicMax(&max_inc[t.i], __float_as_int(inc));
            }
        } else {
            int tpp = 1;
            while (tpp * 2 * nu <= EMD_THREADS) tpp *= 2;
            const int g = tid / tpp, t_in = tid % tpp;
            const bool active = g < nu;
            Top2 t{NEG_BIG, NEG_BIG, -1};
            int j = -1;
            if (active) {
                j = unass[g];
                const float x1 = __ldg(p1 + j * 3 + 0), y1 = __ldg(p1 + j * 3 + 1), z1 = __ldg(p1 + j * 3 + 2);
                for (int k = t_in; k < n; k += tpp) {
                    const float d = bid_value(x1, y1, z1, x2[k], y2[k], z2[k], price[k]);
                    if (d > t.best) { t.better = t.best; t.best = d; t.i = k; }
                    else if (d > t.better) t.better = d;
                }
            }
            // groups are aligned power-of-two lane ranges: butterfly inside the warp (combine is symmetric)
            const int w = tpp < 32 ? tpp : 32;
            for (int off = w >> 1; off > 0; off >>= 1) {
                Top2 o;
                o.best = __shfl_xor_sync(0xffffffffu, t.best, off);
                o.better = __shfl_xor_sync(0xffffffffu, t.better, off);
                o.i = __shfl_xor_sync(0xffffffffu, t.i, off);
                t = combine(t, o);
            }
            if (tpp > 32) {                                             // group spans tpp / 32 whole warps
                if (lane == 0) { red_best[warp] = t.best; red_better[warp] = t.better; red_i[warp] = t.i; }
                __syncthreads();
                if (active && t_in == 0) {
                    for (int q = 1; q < tpp / 32; ++q)
                        t = combine(t, Top2{red_best[warp + q], red_better[warp + q], red_i[warp + q]});
                }
            }
            if (active && t_in == 0) {
                const float inc = __fadd_rn(__fsub_rn(t.best, t.better), eps);
                bid[j] = t.i;
                bid_inc[j] = inc;
                atomicMax(&max_inc[t.i], __float_as_int(inc));
            }
        }
        __syncthreads();
        // ---- GetMax: the highest eligible point index wins a target
        for (int u = tid; u < nu; u += EMD_THREADS) {
            const int j = unass[u];
            const int t = bid[j];
            const double bi = (double)bid_inc[j], mi = (double)__int_as_float(max_inc[t]);
            if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(&max_idx[t], j);
        }
        __syncthreads();
        // ---- Assign
        for (int u = tid; u < nu; u += EMD_THREADS) {
            const int j = unass[u];
            const int t = bid[j];
            if (last || max_idx[t] == j) {
                const int prev = ass_inv[t];
                if (!last && prev != -1) ass[prev] = -1;
                ass_inv[t] = j;
                ass[j] = t;
                price[t] += bid_inc[j];
                max_inc[t] = __float_as_int(NEG_BIG);
            }
        }
        __syncthreads();
    }

    for (int j = tid; j < n; j += EMD_THREADS) {
        const int k = ass[j];
        const float dx = __ldg(p1 + j * 3 + 0) - x2[k], dy = __ldg(p1 + j * 3 + 1) - y2[k], dz = __ldg(p1 + j * 3 + 2) - z2[k];
        dist[(int64_t)blockIdx.x * n + j] = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        assignment[(int64_t)blockIdx.x * n + j] = k;
    }
}

__global__ void emd_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                const float* __restrict__ g, const int32_t* __restrict__ assignment, int64_t total,
                                int n, float* __restrict__ gxyz1) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = (i / n) * n + assignment[i];
        const float gg = g[i] * 2.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) gxyz1[i * 3 + c] = gg * (xyz1[i * 3 + c] - xyz2[k * 3 + c]);
    }
}

inline size_t emd_smem_bytes(int n) { return (size_t)11 * n * 4 + 3 * 32 * 4; }

}  // namespace

extern "C" int spgan_emd_auction(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist,
                                 int32_t* assignment, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(xyz1 && xyz2 && dist && assignment && B >= 0 && n >= 1 && iters >= 1 && eps > 0.f);
    if (B == 0) return SPGAN_OK;
    const size_t smem = emd_smem_bytes(n);
    if (smem > 227 * 1024) return SPGAN_E_UNSUPPORTED;                    // n <= ~5200 points per cloud
    cudaError_t e = cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    emd_auction_kernel<<<B, EMD_THREADS, smem, as_stream(stream)>>>(xyz1, xyz2, n, eps, iters, dist, assignment);
    return spgan_launch_status();
}

extern "C" int spgan_emd_grad(const float* xyz1, const float* xyz2, const float* graddist, const int32_t* assignment,
                              int B, int n, float* gxyz1, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(xyz1 && xyz2 && graddist && assignment && gxyz1 && B >= 0 && n >= 1);
    const int64_t total = (int64_t)B * n;
    if (total == 0) return SPGAN_OK;
    emd_grad_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(xyz1, xyz2, graddist, assignment, total, n, gxyz1);
    return spgan_launch_status();
}
