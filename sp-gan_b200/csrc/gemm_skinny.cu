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

// ---------------------------------------------------------------------------------------------------------------
// Thin products of the first / last layers (Conv1d(3, 64) of critic and generator head, Generator.py:107,
// Discriminator.py:55; Conv1d(64, 3) of the generator tail and its input gradients): C[M,N] = A[M,K] op(B) (+ bias)
// (+ C) with K <= 8 or N <= 4 and M = 1e5..1e6 rows.  These are pure streaming passes (M * (K + N) * 4 bytes); the
// tiled kernels spend their time on tile bookkeeping (25-40 us where 5-6 us of HBM time are needed).  Exact fp32:
// one FMA chain over k = 0..K-1 per output, then + bias, then + C.
constexpr int TK_MAXK = 8, TN_MAXN = 4, TN_MAXK = 256;
constexpr int TNT_ROWS = 512;                     // rows of K per CTA of the thin weight-gradient kernel

// K <= 8: thread = (row, 4 consecutive columns)
template <bool VEC>
__global__ void __launch_bounds__(256)
gemm_thin_k_kernel(int64_t M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                   int64_t ldb, int transB, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                   int accumulate) {
    extern __shared__ float Bs[];                        // [K][Np] (+ bias [Np]), Np = N rounded up to 4
    const int Np = (N + 3) & ~3;
    for (int i = threadIdx.x; i < K * Np; i += blockDim.x) {
        const int kk = i / Np, n = i % Np;
        Bs[i] = n < N ? (transB ? __ldg(B + (int64_t)n * ldb + kk) : __ldg(B + (int64_t)kk * ldb + n)) : 0.f;
    }
    for (int n = threadIdx.x; n < Np; n += blockDim.x) Bs[K * Np + n] = (bias && n < N) ? __ldg(bias + n) : 0.f;
    __syncthreads();
    const int Nq = Np >> 2;
    const int64_t total = M * Nq;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Nq;
        const int n0 = (int)(i - r * Nq) * 4;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* ar = A + r * lda;
        for (int kk = 0; kk < K; ++kk) {
            const float a = __ldg(ar + kk);
            const float4 b = *reinterpret_cast<const float4*>(Bs + kk * Np + n0);
            acc[0] = fmaf(a, b.x, acc[0]); acc[1] = fmaf(a, b.y, acc[1]);
            acc[2] = fmaf(a, b.z, acc[2]); acc[3] = fmaf(a, b.w, acc[3]);
        }
        const float4 bb = *reinterpret_cast<const float4*>(Bs + K * Np + n0);
        float4 o = make_float4(acc[0] + bb.x, acc[1] + bb.y, acc[2] + bb.z, acc[3] + bb.w);
        float* cp = C + r * ldc + n0;
        if (VEC) {
            if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(cp); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *reinterpret_cast<float4*>(cp) = o;
        } else {
            const float ov[4] = {o.x, o.y, o.z, o.w};
            for (int j = 0; j < 4 && n0 + j < N; ++j) cp[j] = accumulate ? cp[j] + ov[j] : ov[j];
        }
    }
}

// N <= 4, K <= 256, K % 4 == 0: thread = row (float4 loads of its A row; W and bias in shared memory)
__global__ void __launch_bounds__(256)
gemm_thin_n_kernel(int64_t M, int N, int K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                   int64_t ldb, int transB, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                   int accumulate) {
    __shared__ float4 Ws[TN_MAXK];                       // Ws[k] = (W[k][0..3]), zero beyond N
    __shared__ float bs[TN_MAXN];
    for (int kk = threadIdx.x; kk < K; kk += blockDim.x) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = 0; n < N; ++n) w[n] = transB ? __ldg(B + (int64_t)n * ldb + kk) : __ldg(B + (int64_t)kk * ldb + n);
        Ws[kk] = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (threadIdx.x < TN_MAXN) bs[threadIdx.x] = (bias && threadIdx.x < N) ? __ldg(bias + threadIdx.x) : 0.f;
    __syncthreads();
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
        const float4* ar = reinterpret_cast<const float4*>(A + r * lda);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k4 = 0; k4 < (K >> 2); ++k4) {
            const float4 a = __ldg(ar + k4);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 w = Ws[4 * k4 + e];
                acc[0] = fmaf(av[e], w.x, acc[0]); acc[1] = fmaf(av[e], w.y, acc[1]);
                acc[2] = fmaf(av[e], w.z, acc[2]); acc[3] = fmaf(av[e], w.w, acc[3]);
            }
        }
        float* cp = C + r * ldc;
        for (int n = 0; n < N; ++n) {
            const float o = acc[n] + bs[n];
            cp[n] = accumulate ? cp[n] + o : o;
        }
    }
}

// Weight gradient of a Conv1d(3, C) (critic conv1, generator head / pc_head): C[Mo,No] (+)= A^T B with A = dY [K, Mo],
// B = X [K, No <= 4], K = 1e5..1e6 points.  Thread = (output row m, k slice): per k one coalesced load of dY[k, :] per
// warp, X[k, 0..3] broadcast, 8 rows in flight per thread; slices folded through shared memory in a fixed order, one
// atomicAdd per output element and CTA (the accumulation contract of the kernel above).
template <int MP>                                  // Mo rounded up to 32 / 64 / 128 / 256 (threads per k slice)
__global__ void __launch_bounds__(256)
gemm_tn_thin_kernel(int Mo, int No, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                    int64_t ldb, float* __restrict__ C, int64_t ldc) {
    constexpr int SL = 256 / MP;                   // k slices per CTA
    __shared__ float4 Bs[TNT_ROWS];                // the CTA's rows of X, padded to 4 columns (zero beyond No / K)
    __shared__ float red[SL][MP][4];
    const int m = threadIdx.x % MP, sl = threadIdx.x / MP;
    const int64_t k0 = (int64_t)blockIdx.x * TNT_ROWS;
    const int rows = (int)((k0 + TNT_ROWS < K ? k0 + TNT_ROWS : K) - k0);
    for (int i = threadIdx.x; i < TNT_ROWS; i += 256) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < rows)
            for (int n = 0; n < No; ++n) v[n] = __ldg(B + (k0 + i) * ldb + n);
        Bs[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = m < Mo;
    const float* ap = A + k0 * lda + m;
    for (int kb = sl * 8; kb < rows; kb += SL * 8) {
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = (kb + u < rows && live) ? __ldg(ap + (int64_t)(kb + u) * lda) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 b = Bs[kb + u < TNT_ROWS ? kb + u : TNT_ROWS - 1];       // (a[u] = 0 beyond `rows`)
            acc[0] = fmaf(a[u], b.x, acc[0]); acc[1] = fmaf(a[u], b.y, acc[1]);
            acc[2] = fmaf(a[u], b.z, acc[2]); acc[3] = fmaf(a[u], b.w, acc[3]);
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) red[sl][m][n] = acc[n];
    __syncthreads();
    if (sl == 0 && live) {
        for (int n = 0; n < No; ++n) {
            float t = red[0][m][n];
            for (int q = 1; q < SL; ++q) t += red[q][m][n];
            atomicAdd(C + (int64_t)m * ldc + n, t);
        }
    }
}

}  // namespace

bool spgan_gemm_thin_supported(int transA, int64_t M, int N, int K, const float* A, int64_t lda) {
    if (transA || M < 4096) return false;
    if (K <= TK_MAXK && N >= 1 && N <= 4096) return true;
    return N <= TN_MAXN && K <= TN_MAXK && (K & 3) == 0 && (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
}

int spgan_gemm_thin(int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                    int64_t ldc, const float* bias, int accumulate, cudaStream_t st) {
    if (K <= TK_MAXK) {
        const int Np = (N + 3) & ~3;
        const size_t smem = (size_t)(K + 1) * Np * sizeof(float);
        const int grid = ew_grid(M * (Np / 4), 256, 16);
        const bool vec = (N & 3) == 0 && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
        if (vec) gemm_thin_k_kernel<true><<<grid, 256, smem, st>>>(M, N, K, A, lda, B, ldb, transB, C, ldc, bias, accumulate);
        else gemm_thin_k_kernel<false><<<grid, 256, smem, st>>>(M, N, K, A, lda, B, ldb, transB, C, ldc, bias, accumulate);
    } else {
        gemm_thin_n_kernel<<<ew_grid(M, 256, 16), 256, 0, st>>>(M, N, K, A, lda, B, ldb, transB, C, ldc, bias, accumulate);
    }
    return spgan_launch_status();
}

bool spgan_gemm_tn_skinny_supported(int64_t Mo, int No, int64_t K) {
    return Mo >= 1 && ((Mo <= 64 && No >= 1 && No <= 32) || (Mo <= 256 && No >= 1 && No <= 4)) && K >= 16384;
}

// C (+)= A^T B; C is zeroed first unless accumulate
int spgan_gemm_tn_skinny(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                         float* C, int64_t ldc, int accumulate, cudaStream_t st) {
    if (!accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)No * sizeof(float), (size_t)Mo, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (No <= 4 && Mo <= 256) {
        const unsigned g = (unsigned)ceil_div64(K, TNT_ROWS);
        if (Mo <= 32) gemm_tn_thin_kernel<32><<<g, 256, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc);
        else if (Mo <= 64) gemm_tn_thin_kernel<64><<<g, 256, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc);
        else if (Mo <= 128) gemm_tn_thin_kernel<128><<<g, 256, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc);
        else gemm_tn_thin_kernel<256><<<g, 256, 0, st>>>((int)Mo, No, K, A, lda, B, ldb, C, ldc);
        return spgan_launch_status();
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
