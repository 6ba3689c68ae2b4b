// fp32 CUDA-core GEMM (engine 0 of spgan_gemm): C[M,N] = op(A) op(B) (+bias) (+C).
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register micro-tiles, register-prefetch double
// buffering.  transA (wgrad: K = number of points) runs split-K with fp32 atomics.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int APAD = 4;

// Loads 4 consecutive elements along the contiguous axis of an operand tile, guarded.
// `r` indexes the non-contiguous axis, `c` the contiguous one.
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t ld, int64_t r, int64_t rmax,
                                        int64_t c, int64_t cmax, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= rmax) return v;
    const float* p = base + r * ld + c;
    if (vec && c + 3 < cmax) {
        v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
        if (c + 0 < cmax) v.x = __ldg(p + 0);
        if (c + 1 < cmax) v.y = __ldg(p + 1);
        if (c + 2 < cmax) v.z = __ldg(p + 2);
        if (c + 3 < cmax) v.w = __ldg(p + 3);
    }
    return v;
}

template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_simt_kernel(int64_t M, int N, int64_t K, const float* __restrict__ A, int64_t lda,
                 const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
                 const float* __restrict__ bias, int accumulate, int64_t k_per_split, bool vecA, bool vecB,
                 bool vecC, bool atomic_out, int64_t c_split_stride) {
    constexpr int TN = BN / 16;           // columns per thread (8 or 4)
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + APAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;

    // ---- global -> register staging assignment
    // A tile: BM x BK.  !TA: A is [M,K], contiguous along k: 4 float4 per row of 16 -> 512 float4, 2 per thread.
    //                    TA: A is [K,M], contiguous along m: 32 float4 per k-row -> 512 float4, 2 per thread.
    // B tile: BN x BK.   TB: B is [N,K], contiguous along k;  !TB: B is [K,N], contiguous along n.
    constexpr int A_V = (BM * BK / 4) / GEMM_THREADS;   // 2
    constexpr int B_V = (BN * BK / 4) / GEMM_THREADS;   // 2 or 1
    float4 ra[A_V], rb[B_V];

    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int v = 0; v < A_V; ++v) {
            const int f = tid + v * GEMM_THREADS;
            if (!TA) {
                const int row = f >> 2, kq = (f & 3) * 4;
                ra[v] = load4(A, lda, m0 + row, M, k0 + kq, kend, vecA);
            } else {
                const int kr = f >> 5, mq = (f & 31) * 4;
                ra[v] = load4(A, lda, k0 + kr, kend, m0 + mq, M, vecA);
            }
        }
#pragma unroll
        for (int v = 0; v < B_V; ++v) {
            const int f = tid + v * GEMM_THREADS;
            if (TB) {
                const int row = f >> 2, kq = (f & 3) * 4;
                rb[v] = load4(B, ldb, n0 + row, N, k0 + kq, kend, vecB);
            } else {
                const int kr = f / (BN / 4), nq = (f % (BN / 4)) * 4;
                rb[v] = load4(B, ldb, k0 + kr, kend, n0 + nq, N, vecB);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int v = 0; v < A_V; ++v) {
            const int f = tid + v * GEMM_THREADS;
            if (!TA) {
                const int row = f >> 2, kq = (f & 3) * 4;
                As[buf][kq + 0][row] = ra[v].x; As[buf][kq + 1][row] = ra[v].y;
                As[buf][kq + 2][row] = ra[v].z; As[buf][kq + 3][row] = ra[v].w;
            } else {
                const int kr = f >> 5, mq = (f & 31) * 4;
                *reinterpret_cast<float4*>(&As[buf][kr][mq]) = ra[v];
            }
        }
#pragma unroll
        for (int v = 0; v < B_V; ++v) {
            const int f = tid + v * GEMM_THREADS;
            if (TB) {
                const int row = f >> 2, kq = (f & 3) * 4;
                Bs[buf][kq + 0][row] = rb[v].x; Bs[buf][kq + 1][row] = rb[v].y;
                Bs[buf][kq + 2][row] = rb[v].z; Bs[buf][kq + 3][row] = rb[v].w;
            } else {
                const int kr = f / (BN / 4), nq = (f % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = rb[v];
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) {
        load_tiles(kbeg);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
            const bool has_next = (k0 + BK) < kend;
            if (has_next) load_tiles(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float bv[TN];
                {
                    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                    if (TN == 8) {
                        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
                        bv[TN - 4] = b1.x; bv[TN - 3] = b1.y; bv[TN - 2] = b1.z; bv[TN - 1] = b1.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (has_next) {
                store_tiles(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

    // ---- epilogue
    C += (int64_t)blockIdx.z * c_split_stride;          // deterministic split-K: every split owns a partial tile
    const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int n = n0 + ((h == 0) ? tx * 4 : BN / 2 + tx * 4);
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            float* cp = C + m * ldc + n;
            if (add_bias) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) v[j] += __ldg(bias + n + j);
            }
            if (atomic_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) atomicAdd(cp + j, v[j]);
            } else if (vecC && n + 3 < N) {
                if (accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(cp);
                    v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
                }
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) cp[j] = accumulate ? cp[j] + v[j] : v[j];
            }
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// out[m,n] = sum_z partial[z][m][n] (+ bias[n]) (+ out[m,n]): fixed summation order => deterministic
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int N,
                                     float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                     int accumulate) {
    const int64_t total = M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N;
        const int n = (int)(i - m * N);
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += __ldg(partial + (int64_t)z * total + i);
        if (bias) s += __ldg(bias + n);
        float* cp = C + m * ldc + n;
        *cp = accumulate ? *cp + s : s;
    }
}

constexpr int kSmallMSplits = 16;
constexpr int kSmallMRows = 128;

template <int BN>
int launch_simt(int transA, int transB, int64_t M, int N, int64_t K, const float* A, int64_t lda, const float* B,
                int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate, cudaStream_t st,
                void* workspace, size_t workspace_bytes) {
    const int64_t mt = ceil_div64(M, BM);
    const int nt = (N + BN - 1) / BN;
    // small-batch shapes (the critic's MLP on B rows, the generator's global MLP): a handful of CTAs would walk
    // K serially.  Split K over up to 16 CTAs per tile into private partial tiles, then reduce in a fixed order.
    if (mt == 1 && M <= kSmallMRows && K >= 256 && K < 2048 && workspace != nullptr) {
        int splits = (int)(K / 64);
        if (splits > kSmallMSplits) splits = kSmallMSplits;
        const size_t need = (size_t)splits * M * N * sizeof(float);
        if (splits > 1 && workspace_bytes >= need) {
            float* partial = reinterpret_cast<float*>(workspace);
            const int64_t kps = ceil_div64(ceil_div64(K, splits), BK) * BK;
            splits = (int)ceil_div64(K, kps);
            const bool vecA = (lda % 4 == 0) && aligned16(A);
            const bool vecB = (ldb % 4 == 0) && aligned16(B);
            const bool vecP = (N % 4 == 0) && aligned16(partial);
            dim3 grid(1, (unsigned)nt, (unsigned)splits);
#define SPGAN_LAUNCH_P(TA, TB)                                                                                 \
    gemm_simt_kernel<BN, TA, TB><<<grid, GEMM_THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, partial, N, nullptr, 0, kps, \
                                                                 vecA, vecB, vecP, false, M * (int64_t)N)
            if (!transA && !transB) SPGAN_LAUNCH_P(false, false);
            else if (!transA && transB) SPGAN_LAUNCH_P(false, true);
            else if (transA && !transB) SPGAN_LAUNCH_P(true, false);
            else SPGAN_LAUNCH_P(true, true);
#undef SPGAN_LAUNCH_P
            splitk_reduce_kernel<<<ew_grid(M * N, 256), 256, 0, st>>>(partial, splits, M, N, C, ldc, bias, accumulate);
            return spgan_launch_status();
        }
    }
    // split-K (fp32 atomics) when the output grid cannot fill the machine and K is long: weight gradients,
    // K = #points.  Forward-path shapes (K <= 1280) never split, so forward results are run-to-run
    // deterministic.
    int64_t splits = 1;
    const int64_t tiles = mt * nt;
    if (tiles < 2 * kNumSMs && K >= 2048) {
        splits = (4 * kNumSMs + tiles - 1) / tiles;
        const int64_t max_splits = K / 512;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        if (splits > 65535) splits = 65535;
    }
    int64_t kps = ceil_div64(ceil_div64(K, splits), BK) * BK;
    splits = ceil_div64(K, kps);
    const bool atomic_out = splits > 1;
    if (atomic_out && !accumulate) {
        // zero the output rows (ldc may exceed N)
        cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st);
        if (e != cudaSuccess) return (int)e;
    }
    const bool vecA = (lda % 4 == 0) && aligned16(A);
    const bool vecB = (ldb % 4 == 0) && aligned16(B);
    const bool vecC = (ldc % 4 == 0) && aligned16(C);
    if (mt > 0x7fffffffLL || nt > 65535) return SPGAN_E_UNSUPPORTED;
    dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)splits);
#define SPGAN_LAUNCH(TA, TB)                                                                              \
    gemm_simt_kernel<BN, TA, TB><<<grid, GEMM_THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, \
                                                                 accumulate, kps, vecA, vecB, vecC, atomic_out, 0)
    if (!transA && !transB) SPGAN_LAUNCH(false, false);
    else if (!transA && transB) SPGAN_LAUNCH(false, true);
    else if (transA && !transB) SPGAN_LAUNCH(true, false);
    else SPGAN_LAUNCH(true, true);
#undef SPGAN_LAUNCH
    return spgan_launch_status();
}

}  // namespace

int spgan_gemm_simt(int transA, int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                    int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate, cudaStream_t st,
                    void* workspace, size_t workspace_bytes) {
    // narrow tiles also when 128-wide ones would leave most SMs idle (the M = 64 heads and their weight gradients:
    // 32 CTAs of 128 x 128 -> 64 of 128 x 64); the k order of every output's FMA chain does not depend on the tile width
    if (N <= 64 || ceil_div64(M, BM) * ((N + 127) / 128) < kNumSMs / 2)
        return launch_simt<64>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, st, workspace,
                               workspace_bytes);
    return launch_simt<128>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, st, workspace,
                            workspace_bytes);
}


size_t spgan_gemm_tc_workspace(int N, int K);
bool spgan_gemm_tc_supported(int transA, int64_t M, int N, int K);
int spgan_gemm_tc(int mode, int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate, void* workspace,
                  cudaStream_t st);

size_t spgan_gemm_ts_workspace(int N, int K);
bool spgan_gemm_ts_supported(int64_t M, int N, int K, const float* A, int64_t lda);
int spgan_gemm_ts(int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, const float* bias, int accumulate, const float* a_scale, const float* a_shift, float a_slope,
                  float* col_sum, float* col_sqsum, void* workspace, cudaStream_t st);
bool spgan_gemm_tsk_supported(int64_t M, int N, int K, const float* A, int64_t lda);
bool spgan_gemm_wg_supported(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb);
size_t spgan_gemm_wg_workspace(int64_t Mo, int No, int64_t K);
int spgan_gemm_wg(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, int accumulate, void* workspace, cudaStream_t st, const float* b_scale, const float* b_shift,
                  float b_slope);
static bool wg_enabled() {
    static const bool on = [] { const char* e = getenv("SPGAN_WG"); return !(e && e[0] == '0'); }();
    return on;
}
static bool ts_enabled() {
    static const bool on = [] { const char* e = getenv("SPGAN_TS"); return !(e && e[0] == '0'); }();
    return on;
}

bool spgan_gemm_tc_tn_supported(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B,
                                int64_t ldb);
int spgan_gemm_tc_tn(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                     int64_t ldc, int accumulate, void* workspace, cudaStream_t st);

bool spgan_gemm_thin_supported(int transA, int64_t M, int N, int K, const float* A, int64_t lda);
int spgan_gemm_thin(int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                    int64_t ldc, const float* bias, int accumulate, cudaStream_t st);
bool spgan_gemm_tn_skinny_supported(int64_t Mo, int No, int64_t K);
int spgan_gemm_tn_skinny(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                         float* C, int64_t ldc, int accumulate, cudaStream_t st);

extern "C" size_t spgan_gemm_workspace(int engine, int N, int K) {
    if (N < 1 || K < 1) return 0;
    // deterministic split-K partial tiles of the small-batch path (any engine)
    const size_t small_m = (K >= 256 && K < 2048) ? (size_t)kSmallMSplits * kSmallMRows * N * sizeof(float) : 0;
    size_t tc = (engine >= 1 && engine <= 3) ? spgan_gemm_tc_workspace(N, K) : 0;
    if (engine == 3) { const size_t ts = spgan_gemm_ts_workspace(N, K); if (ts > tc) tc = ts; }
    return small_m > tc ? small_m : tc;
}

/* 1 when spgan_gemm(engine 3, no transA) sends this product to the chunked-K TMEM-resident kernel (gemm_ts.cu): the
 * layout of the split weight in the workspace depends on the route, so a caller that caches workspaces keys on it */
extern "C" size_t spgan_gemm_bigk_route(int64_t M, int N, int K, const float* A, int64_t lda) {
    return (ts_enabled() && spgan_gemm_tsk_supported(M, N, K, A, lda)) ? 1 : 0;
}

extern "C" int spgan_gemm(int transA, int transB, int64_t M, int N, int K, const float* A, int64_t lda,
                          const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate,
                          int engine, void* workspace, size_t workspace_bytes, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(A && B && C && M >= 0 && N >= 1 && K >= 1);
    const int presplit = transB & 2;     // tensor engines only: the workspace holds the split weight of an earlier call
    transB &= 1;
    SPGAN_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N);
    if (M == 0) return SPGAN_OK;
    // engine 3, K <= 256: the TMEM-resident-A kernel (gemm_ts.cu); SPGAN_TS=0 routes these to gemm_tc.cu instead
    if (engine == 3 && !transA && workspace != nullptr && ts_enabled() && spgan_gemm_ts_supported(M, N, K, A, lda) &&
        workspace_bytes >= spgan_gemm_ts_workspace(N, K) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0)
        return spgan_gemm_ts(transB | presplit, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, nullptr, nullptr, 1.f, nullptr,
                             nullptr, workspace, as_stream(stream));
    // engine 3, K > 256, K % 128 == 0: the same kernel family with K walked in chunks (gemm_tsk_kernel)
    if (engine == 3 && !transA && workspace != nullptr && ts_enabled() && spgan_gemm_tsk_supported(M, N, K, A, lda) &&
        workspace_bytes >= spgan_gemm_ts_workspace(N, K) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0)
        return spgan_gemm_ts(transB | presplit, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, nullptr, nullptr, 1.f, nullptr,
                             nullptr, workspace, as_stream(stream));
    // tcgen05 engines: 1 = TF32x3, 2 = BF16x3, 3 = FP16Sx3 (gemm_tc.cu) for the forward / dgrad products
    if (engine >= 1 && engine <= 3 && workspace != nullptr && spgan_gemm_tc_supported(transA, M, N, K) &&
        workspace_bytes >= spgan_gemm_tc_workspace(N, K) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0)
        return spgan_gemm_tc(engine - 1, transB | presplit, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, workspace,
                             as_stream(stream));
    // weight gradients: C[M,N] = A^T B with A [K,M], B [K,N], K = #points.  Engine 3: the TMA / TMEM kernel with
    // deterministic split-K partials (gemm_wg.cu) when the caller's workspace holds them; SPGAN_WG=0 disables it
    if (engine == 3 && workspace != nullptr && transA && !transB && bias == nullptr && wg_enabled() &&
        (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && spgan_gemm_wg_supported(M, N, K, A, lda, B, ldb) &&
        workspace_bytes >= spgan_gemm_wg_workspace(M, N, K))
        return spgan_gemm_wg(M, N, K, A, lda, B, ldb, C, ldc, accumulate, workspace, as_stream(stream), nullptr, nullptr, 1.f);
    if (engine == 3) engine = 1;          // otherwise the TF32x3 transposing kernel for every tensor engine
    // weight gradients, first generation (TF32x3, atomic flush)
    if ((engine == 1 || engine == 2) && workspace != nullptr && workspace_bytes >= 256 && transA && !transB &&
        bias == nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
        spgan_gemm_tc_tn_supported(M, N, K, A, lda, B, ldb))
        return spgan_gemm_tc_tn(M, N, K, A, lda, B, ldb, C, ldc, accumulate, workspace, as_stream(stream));
    // tall-skinny weight gradients (tiny output, K = #points / #edges): stream K once (any engine)
    if (transA && !transB && bias == nullptr && spgan_gemm_tn_skinny_supported(M, N, K))
        return spgan_gemm_tn_skinny(M, N, K, A, lda, B, ldb, C, ldc, accumulate, as_stream(stream));
    // thin products (K <= 8 or N <= 4 over >= 4096 rows): streaming kernels, exact fp32 (any engine)
    if (spgan_gemm_thin_supported(transA, M, N, K, A, lda))
        return spgan_gemm_thin(transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, as_stream(stream));
    return spgan_gemm_simt(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, as_stream(stream),
                           workspace, workspace_bytes);
}

extern "C" size_t spgan_gemm_fused_workspace(int64_t M, int N, int K, const float* A, int64_t lda) {
    if (N < 1 || K < 1 || !spgan_gemm_ts_supported(M, N, K, A, lda)) return 0;
    return spgan_gemm_ts_workspace(N, K);
}

extern "C" size_t spgan_gemm_fused_stats_rows(int64_t M) {
    if (M <= 0) return 0;
    const int64_t tiles = (M + 127) / 128;
    return (size_t)(4 * (tiles < kNumSMs ? tiles : kNumSMs));           // one partial row per (CTA, lane quarter)
}

extern "C" int spgan_gemm_fused(int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                                int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate,
                                const float* a_scale, const float* a_shift, float a_slope, float* col_sum,
                                float* col_sqsum, void* workspace, size_t workspace_bytes, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(A && B && C && workspace && M >= 1 && N >= 1 && K >= 1);
    SPGAN_CHECK_ARG(lda >= K && ldb >= ((transB & 1) ? K : N) && ldc >= N);
    SPGAN_CHECK_ARG((a_scale == nullptr) == (a_shift == nullptr) && (col_sum == nullptr) == (col_sqsum == nullptr));
    if (col_sum != nullptr && N > 256) return SPGAN_E_UNSUPPORTED;     // column statistics: up to 4 column tiles
    if (!spgan_gemm_ts_supported(M, N, K, A, lda) || workspace_bytes < spgan_gemm_ts_workspace(N, K) ||
        (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return SPGAN_E_UNSUPPORTED;
    return spgan_gemm_ts(transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, a_scale, a_shift, a_slope, col_sum,
                         col_sqsum, workspace, as_stream(stream));
}

extern "C" size_t spgan_gemm_wgrad_workspace(int64_t Mo, int No, int64_t K) {
    if (Mo < 1 || No < 1 || K < 1) return 256;
    return spgan_gemm_wg_workspace(Mo, No, K);
}

extern "C" int spgan_gemm_wgrad_fused(int64_t Mo, int No, int64_t K, const float* dY, int64_t ldy, const float* X,
                                      int64_t ldx, const float* x_scale, const float* x_shift, float x_slope, float* C,
                                      int64_t ldc, int accumulate, void* workspace, size_t workspace_bytes,
                                      spgan_stream_t stream) {
    SPGAN_CHECK_ARG(dY && X && C && Mo >= 1 && No >= 1 && K >= 1 && ldy >= Mo && ldx >= No && ldc >= No);
    SPGAN_CHECK_ARG((x_scale == nullptr) == (x_shift == nullptr) && x_slope > 0.f && x_slope <= 1.f);
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0 || !wg_enabled() ||
        !spgan_gemm_wg_supported(Mo, No, K, dY, ldy, X, ldx) || workspace_bytes < spgan_gemm_wg_workspace(Mo, No, K))
        return SPGAN_E_UNSUPPORTED;
    return spgan_gemm_wg(Mo, No, K, dY, ldy, X, ldx, C, ldc, accumulate, workspace, as_stream(stream), x_scale, x_shift,
                         x_slope);
}
