// Tall-skinny weight gradients C[Mo,No] (+)= A^T B with A = dY [K, Mo], B = X [K, No], K = number of points or
// edges (1e5 .. 1e6) and a tiny output (Mo <= 64, No <= 32): EdgeConv1's conv_w.3 (64 x 32, K = 1.3 M edges) and the
// first layers of the critic / generator (64 x 3, K = 131 072).  The general kernels waste 4-20x of their 128 x BN
// tiles on these shapes; here a CTA streams a contiguous K range through shared memory (register-staged
// prefetch of the next 64 rows), every thread owns a TM x TN block of the whole output, and the per-CTA partial
// is flushed with fp32 atomics (same accumulation contract as the split-K path it replaces).  HBM-bound:
// K * (Mo + No) * 4 bytes are read exactly once.
#include "common.cuh"

namespace {

constexpr int SK_THREADS = 256;
constexpr int KC = 64;                       // rows per staged chunk

template <int TM, int TN>
__global__ void __launch_bounds__(SK_THREADS)
gemm_tn_skinny_kernel(int Mo, int No, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                      int64_t ldb, float* __restrict__ C, int64_t ldc, int64_t rows_per_cta) {
    constexpr int MP = 16 * TM, NP = 16 * TN;            // padded output extent covered by the 16 x 16 thread grid
    __shared__ __align__(16) float As[KC][MP];
    __shared__ __align__(16) float Bs[KC][NP];
    constexpr int A_PER = KC * MP / SK_THREADS, B_PER = (KC * NP + SK_THREADS - 1) / SK_THREADS;
    const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
    const int64_t k0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t k1 = (k0 + rows_per_cta < K) ? k0 + rows_per_cta : K;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[A_PER], rb[B_PER];
    auto load = [&](int64_t kb) {
#pragma unroll
        for (int v = 0; v < A_PER; ++v) {
            const int e = tid + v * SK_THREADS, r = e / MP, c = e % MP;
            ra[v] = (kb + r < k1 && c < Mo) ? __ldg(A + (kb + r) * lda + c) : 0.f;
        }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) {
            const int e = tid + v * SK_THREADS, r = e / NP, c = e % NP;
            rb[v] = (e < KC * NP && kb + r < k1 && c < No) ? __ldg(B + (kb + r) * ldb + c) : 0.f;
        }
    };
    if (k0 < k1) load(k0);
    for (int64_t kb = k0; kb < k1; kb += KC) {
        __syncthreads();                                   // previous chunk consumed
#pragma unroll
        for (int v = 0; v < A_PER; ++v) { const int e = tid + v * SK_THREADS; As[e / MP][e % MP] = ra[v]; }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) { const int e = tid + v * SK_THREADS; if (e < KC * NP) Bs[e / NP][e % NP] = rb[v]; }
        __syncthreads();
        if (kb + KC < k1) load(kb + KC);                   // next chunk's loads fly during the FMAs
#pragma unroll 8
        for (int kk = 0; kk < KC; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][tm * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tn * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int m = tm * TM + i, n = tn * TN + j;
            if (m < Mo && n < No) atomicAdd(C + (int64_t)m * ldc + n, acc[i][j]);
        }
}

}  // namespace

bool spgan_gemm_tn_skinny_supported(int64_t Mo, int No, int64_t K) {
    return Mo >= 1 && Mo <= 64 && No >= 1 && No <= 32 && K >= 16384;
}

// C (+)= A^T B; C is zeroed first unless accumulate
int spgan_gemm_tn_skinny(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                         float* C, int64_t ldc, int accumulate, cudaStream_t st) {
    if (!accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)No * sizeof(float), (size_t)Mo, st);
        if (e != cudaSuccess) return (int)e;
    }
    int64_t ctas = 2 * kNumSMs;
    int64_t rows = ceil_div64(ceil_div64(K, ctas), KC) * KC;
    if (rows < 4 * KC) rows = 4 * KC;
    ctas = ceil_div64(K, rows);
    if (No <= 16)
        gemm_tn_skinny_kernel<4, 1><<<(unsigned)ctas, SK_THREADS, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc, rows);
    else
        gemm_tn_skinny_kernel<4, 2><<<(unsigned)ctas, SK_THREADS, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc, rows);
    return spgan_launch_status();
}
