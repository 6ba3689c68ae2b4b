// kNN graph build + neighbour grouping, bit-exact against the reference's CPU arithmetic.
// Replaces Generation/modules.py:695-720.  See include/spgan_b200.h for the contract and
// DESIGN.md ("kNN kernel") for the tiling.
#include "common.cuh"
#include <float.h>

namespace {

// ---------------------------------------------------------------------------------------
// squared norms: products rounded first (the reference materialises x**2), cascade order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int ceil_log2_dev(long long x) {
    if (x <= 2) return 1;
    return 64 - __clzll((unsigned long long)(x - 1));
}

__device__ float sqnorm_cascade(const float* __restrict__ col, int C, int64_t stride) {
    int p = ceil_log2_dev(C) / 4;
    if (p < 4) p = 4;
    const int step = 1 << p;
    const int mask = step - 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    while (i + step <= C) {
        for (int j = 0; j < step; ++j, ++i) {
            const float v = col[(int64_t)i * stride];
            acc[0] = __fadd_rn(acc[0], __fmul_rn(v, v));
        }
#pragma unroll
        for (int l = 1; l < 4; ++l) {
            acc[l] = __fadd_rn(acc[l], acc[l - 1]);
            acc[l - 1] = 0.f;
            if ((i & (mask << (l * p))) != 0) break;
        }
    }
    for (; i < C; ++i) {
        const float v = col[(int64_t)i * stride];
        acc[0] = __fadd_rn(acc[0], __fmul_rn(v, v));
    }
#pragma unroll
    for (int l = 1; l < 4; ++l) acc[0] = __fadd_rn(acc[0], acc[l]);
    return acc[0];
}

__device__ float sqnorm_ilp4(const float* __restrict__ col, int C, int64_t stride) {
    const int G = C / 4;
    int p = ceil_log2_dev(G) / 4;
    if (p < 4) p = 4;
    const int step = 1 << p;
    const int mask = step - 1;
    float acc[4][4];
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
    int i = 0;
    while (i + step <= G) {
        for (int j = 0; j < step; ++j, ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float v = col[(int64_t)(i * 4 + q) * stride];
                acc[0][q] = __fadd_rn(acc[0][q], __fmul_rn(v, v));
            }
        }
#pragma unroll
        for (int l = 1; l < 4; ++l) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[l][q] = __fadd_rn(acc[l][q], acc[l - 1][q]);
                acc[l - 1][q] = 0.f;
            }
            if ((i & (mask << (l * p))) != 0) break;
        }
    }
    for (; i < G; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float v = col[(int64_t)(i * 4 + q) * stride];
            acc[0][q] = __fadd_rn(acc[0][q], __fmul_rn(v, v));
        }
    }
#pragma unroll
    for (int l = 1; l < 4; ++l)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[0][q] = __fadd_rn(acc[0][q], acc[l][q]);
    for (int c = G * 4; c < C; ++c) {
        const float v = col[(int64_t)c * stride];
        acc[0][0] = __fadd_rn(acc[0][0], __fmul_rn(v, v));
    }
#pragma unroll
    for (int q = 1; q < 4; ++q) acc[0][0] = __fadd_rn(acc[0][0], acc[0][q]);
    return acc[0][0];
}

__global__ void sqnorm_kernel(const float* __restrict__ x, int B, int C, int N, int main_cols,
                              float* __restrict__ xs) {
    const int64_t total = (int64_t)B * N;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / N), n = (int)(t % N);
        const float* col = x + (int64_t)b * C * N + n;
        xs[t] = (n < main_cols) ? sqnorm_cascade(col, C, N) : sqnorm_ilp4(col, C, N);
    }
}

// ---------------------------------------------------------------------------------------
// fused kNN (+ group) kernel
//   CTA = 64 query points of one cloud; loops over tiles of 128 candidate points.
//   Phase A (all 256 threads): register-tiled 4x8 fp32 FMA chains over channels, chunks of 32
//     channels staged in shared memory; chain order c = 0..C-1 exactly as the CPU bmm.
//   Phase B (per warp, 8 queries each): distances go through a shared tile; every lane scans 4
//     candidates per query against the running k-th best; survivors are inserted into a sorted
//     (dist, index) list distributed over the lanes of the warp (lane r holds rank r).
//   Epilogue: ranks 1..k -> idx; optional grouped edge features.
// ---------------------------------------------------------------------------------------
constexpr int QT = 64;       // queries per CTA
constexpr int CT = 128;      // candidates per tile
constexpr int CK = 32;       // channels per staged chunk
constexpr int KNN_THREADS = 256;
constexpr int DPAD = 4;

struct KnnSmem {
    float q[CK][QT];            //  8 KB
    float c[CK][CT];            // 16 KB
    float d[QT][CT + DPAD];     // 33 KB
    float xs_q[QT];
    float xs_c[CT];
};

__device__ __forceinline__ bool lex_less(float d0, int j0, float d1, int j1) {
    return d0 < d1 || (d0 == d1 && j0 < j1);
}
// Selection key of a distance.  NaN and +inf (a diverged generator step feeds non-finite features into
// EdgeConv2's graph) become FLT_MAX: still ordered by candidate index, and still lex_less than the list's
// (FLT_MAX, 0x7fffffff) sentinel, so every list entry ends up a VALID candidate index as with torch.sort
// (which puts NaN last).  Finite distances are untouched: one FMNMX per candidate.
__device__ __forceinline__ float ord_key(float d) { return fminf(d, FLT_MAX); }
// A list still holding the sentinel (fewer than k+1 candidates) never reaches memory as an index
__device__ __forceinline__ int valid_nbr(int j, int self, int N) { return j < N ? j : min(self, N - 1); }

__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_group_kernel(const float* __restrict__ x, const float* __restrict__ xs, int B, int C, int N, int k,
                 int32_t* __restrict__ idx, float* __restrict__ ee) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KnnSmem& s = *reinterpret_cast<KnnSmem*>(smem_raw);
    __shared__ int nbr[QT][32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;     // 16 x 16 thread grid: ty -> 4 queries, tx -> 8 candidates
    const int q_tiles = (N + QT - 1) / QT;
    const int b = blockIdx.x / q_tiles;
    const int i0 = (blockIdx.x % q_tiles) * QT;
    const float* xb = x + (int64_t)b * C * N;
    const float* xsb = xs + (int64_t)b * N;
    const int K1 = k + 1;

    if (tid < QT) s.xs_q[tid] = (i0 + tid < N) ? xsb[i0 + tid] : 0.f;

    // sorted lists of this warp's 8 queries: lane r holds rank r
    float ld[8];
    int lj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { ld[u] = FLT_MAX; lj[u] = 0x7fffffff; }
    // (FLT_MAX, 0x7fffffff) is the sentinel: every real candidate, ord_key'ed, sorts before it

    for (int j0 = 0; j0 < N; j0 += CT) {
        float acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;

        for (int c0 = 0; c0 < C; c0 += CK) {
            __syncthreads();          // previous chunk (and previous tile's phase B) fully consumed
            // stage chunk: q[cc][0..63] and c[cc][0..127]; rows beyond C are not read later
            for (int e = tid; e < CK * QT; e += KNN_THREADS) {
                const int cc = e / QT, qq = e % QT;
                const int ch = c0 + cc, i = i0 + qq;
                s.q[cc][qq] = (ch < C && i < N) ? __ldg(xb + (int64_t)ch * N + i) : 0.f;
            }
            for (int e = tid; e < CK * CT; e += KNN_THREADS) {
                const int cc = e / CT, jj = e % CT;
                const int ch = c0 + cc, j = j0 + jj;
                s.c[cc][jj] = (ch < C && j < N) ? __ldg(xb + (int64_t)ch * N + j) : 0.f;
            }
            if (c0 == 0 && tid < CT) s.xs_c[tid] = (j0 + tid < N) ? xsb[j0 + tid] : 0.f;
            __syncthreads();
            const int cmax = min(CK, C - c0);
#pragma unroll 4
            for (int cc = 0; cc < cmax; ++cc) {
                const float4 qv = *reinterpret_cast<const float4*>(&s.q[cc][ty * 4]);
                const float4 c0v = *reinterpret_cast<const float4*>(&s.c[cc][tx * 4]);
                const float4 c1v = *reinterpret_cast<const float4*>(&s.c[cc][64 + tx * 4]);
                const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
                const float ca[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[a][c] = __fmaf_rn(qa[a], ca[c], acc[a][c]);
            }
        }
        // dist = (-2*dot + xs_i) + xs_j, two separate roundings (modules.py:696,699)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float xq = s.xs_q[ty * 4 + a];
            float out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int jj = (c < 4) ? tx * 4 + c : 64 + tx * 4 + (c - 4);
                out[c] = ord_key(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, acc[a][c]), xq), s.xs_c[jj]));
            }
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][tx * 4]) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][64 + tx * 4]) = make_float4(out[4], out[5], out[6], out[7]);
        }
        __syncthreads();

        // ---- phase B: selection, warp `warp` owns queries warp*8 .. warp*8+7
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int qq = warp * 8 + u;
            const float4 dv = *reinterpret_cast<const float4*>(&s.d[qq][lane * 4]);
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
            float tau = __shfl_sync(0xffffffffu, ld[u], K1 - 1);
            int tauj = __shfl_sync(0xffffffffu, lj[u], K1 - 1);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = j0 + lane * 4 + t;
                const float d = dd[t];
                bool pass = (j < N) && lex_less(d, j, tau, tauj);
                unsigned m = __ballot_sync(0xffffffffu, pass);
                while (m) {
                    const int src = __ffs(m) - 1;
                    const float nd = __shfl_sync(0xffffffffu, d, src);
                    const int nj = __shfl_sync(0xffffffffu, j, src);
                    // position = number of list entries strictly before (nd, nj)
                    const bool before = (lane < K1) && lex_less(ld[u], lj[u], nd, nj);
                    const int pos = __popc(__ballot_sync(0xffffffffu, before));
                    const float upd = __shfl_up_sync(0xffffffffu, ld[u], 1);
                    const int upj = __shfl_up_sync(0xffffffffu, lj[u], 1);
                    if (lane == pos) { ld[u] = nd; lj[u] = nj; }
                    else if (lane > pos && lane < K1) { ld[u] = upd; lj[u] = upj; }
                    tau = __shfl_sync(0xffffffffu, ld[u], K1 - 1);
                    tauj = __shfl_sync(0xffffffffu, lj[u], K1 - 1);
                    pass = pass && (lane != src) && lex_less(d, j, tau, tauj);
                    m = __ballot_sync(0xffffffffu, pass);
                }
            }
        }
    }

    // ---- epilogue: ranks 1..k
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int qq = warp * 8 + u;
        const int i = i0 + qq;
        if (lane >= 1 && lane < K1) {
            const int jn = valid_nbr(lj[u], i, N);
            nbr[qq][lane - 1] = jn;
            if (i < N) idx[((int64_t)b * N + i) * k + (lane - 1)] = jn;
        }
    }
    if (ee == nullptr) return;
    __syncthreads();
    const int nq = min(QT, N - i0);
    const int per_ch = nq * k;
    for (int c = 0; c < C; ++c) {
        const float* row = xb + (int64_t)c * N;
        float* e_ctr = ee + (((int64_t)b * 2 * C + c) * N + i0) * k;
        float* e_dif = ee + (((int64_t)b * 2 * C + C + c) * N + i0) * k;
        for (int e = tid; e < per_ch; e += KNN_THREADS) {
            const int qq = e / k, r = e - qq * k;
            const float ctr = __ldg(row + i0 + qq);
            const float nb = __ldg(row + nbr[qq][r]);
            e_ctr[e] = ctr;
            e_dif[e] = nb - ctr;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fast variant for N % 4 == 0 and C <= 128 (every shape on the generator's path): the 64-query tile
// stays resident in shared memory for the whole kernel, candidate chunks are double buffered with
// cp.async (loads of chunk i+1 overlap the FMA chains of chunk i), and the selection phase rejects a
// whole 4-candidate group per lane with one ballot in the common case.  Arithmetic and tie order are
// identical to knn_group_kernel (same FMA chains, same (dist, index) lists).
// ---------------------------------------------------------------------------------------
constexpr int QMAXC = 128;

struct KnnFastSmem {
    float q[QMAXC][QT];          // 32 KB, channel-major, resident
    float c[2][CK][CT];          // 32 KB, double buffered candidate chunk
    float d[QT][CT + DPAD];      // 33 KB
    float xs_q[QT];
    float xs_c[2][CT];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;            // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_group_fast_kernel(const float* __restrict__ x, const float* __restrict__ xs, int B, int C, int N, int k,
                      int32_t* __restrict__ idx, float* __restrict__ ee) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KnnFastSmem& s = *reinterpret_cast<KnnFastSmem*>(smem_raw);
    __shared__ int nbr[QT][32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int q_tiles = (N + QT - 1) / QT;
    const int b = blockIdx.x / q_tiles;
    const int i0 = (blockIdx.x % q_tiles) * QT;
    const float* xb = x + (int64_t)b * C * N;
    const float* xsb = xs + (int64_t)b * N;
    const int K1 = k + 1;
    const int n_chunks = (C + CK - 1) / CK;
    const int n_tiles = (N + CT - 1) / CT;
    const int total = n_tiles * n_chunks;

    // resident query tile (zero beyond N / C)
    for (int e = tid; e < C * (QT / 4); e += KNN_THREADS) {
        const int ch = e / (QT / 4), q4 = (e % (QT / 4)) * 4;
        cp_async16(&s.q[ch][q4], xb + (int64_t)ch * N + i0 + q4, i0 + q4 < N);
    }
    if (tid < QT) s.xs_q[tid] = (i0 + tid < N) ? xsb[i0 + tid] : 0.f;
    auto stage = [&](int it, int buf) {          // it -> (tile, chunk)
        const int tile = it / n_chunks, chunk = it % n_chunks;
        const int j0 = tile * CT, c0 = chunk * CK;
        for (int e = tid; e < CK * (CT / 4); e += KNN_THREADS) {
            const int cc = e / (CT / 4), j4 = (e % (CT / 4)) * 4;
            const int ch = c0 + cc;
            cp_async16(&s.c[buf][cc][j4], xb + (int64_t)(ch < C ? ch : 0) * N + j0 + j4, ch < C && j0 + j4 < N);
        }
        // candidates beyond N get |x_j|^2 = +inf => dist = +inf: they never pass the selection filter, which
        // therefore needs no j < N test
        if (chunk == 0 && tid < CT) s.xs_c[tile & 1][tid] = (j0 + tid < N) ? xsb[j0 + tid] : INFINITY;
    };
    stage(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    float ld[8];
    int lj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { ld[u] = FLT_MAX; lj[u] = 0x7fffffff; }

    float acc[4][8];
    for (int it = 0; it < total; ++it) {
        const int tile = it / n_chunks, chunk = it % n_chunks;
        const int buf = it & 1;
        if (chunk == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                          // chunk `it` visible; everyone is done with buffer buf^1
        if (it + 1 < total) stage(it + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");

        const int c0 = chunk * CK;
        const int cmax = min(CK, C - c0);
#pragma unroll 4
        for (int cc = 0; cc < cmax; ++cc) {
            const float4 qv = *reinterpret_cast<const float4*>(&s.q[c0 + cc][ty * 4]);
            const float4 c0v = *reinterpret_cast<const float4*>(&s.c[buf][cc][tx * 4]);
            const float4 c1v = *reinterpret_cast<const float4*>(&s.c[buf][cc][64 + tx * 4]);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
            const float ca[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[a][c] = __fmaf_rn(qa[a], ca[c], acc[a][c]);
        }
        if (chunk != n_chunks - 1) continue;

        // ---- tile complete: distances -> shared, then selection
        const int j0 = tile * CT;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float xq = s.xs_q[ty * 4 + a];
            float out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int jj = (c < 4) ? tx * 4 + c : 64 + tx * 4 + (c - 4);
                out[c] = ord_key(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, acc[a][c]), xq), s.xs_c[tile & 1][jj]));
            }
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][tx * 4]) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][64 + tx * 4]) = make_float4(out[4], out[5], out[6], out[7]);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int qq = warp * 8 + u;
            const float4 dv = *reinterpret_cast<const float4*>(&s.d[qq][lane * 4]);
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
            const int jb = j0 + lane * 4;
            if (tile == 0) {
                // bulk initialisation (N >= 128: all 128 candidates of tile 0 are valid): a warp bitonic sort of
                // one candidate per lane by (dist, index) puts rank r into lane r -- 15 exchange steps instead of
                // ~25 serial insertions; the other 96 candidates then go through the normal filter.
                float sd = dd[0];
                int sj = jb;
#pragma unroll
                for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                    for (int st = k2 >> 1; st > 0; st >>= 1) {
                        const float od = __shfl_xor_sync(0xffffffffu, sd, st);
                        const int oj = __shfl_xor_sync(0xffffffffu, sj, st);
                        const bool keep_min = ((lane & st) == 0) == ((lane & k2) == 0);
                        const bool other_less = lex_less(od, oj, sd, sj);       // keys are distinct (distinct j)
                        if (keep_min == other_less) { sd = od; sj = oj; }
                    }
                }
                ld[u] = sd; lj[u] = sj;          // lanes >= K1 hold entries that are never read or written back
            }
            float tau = __shfl_sync(0xffffffffu, ld[u], K1 - 1);
            int tauj = __shfl_sync(0xffffffffu, lj[u], K1 - 1);
            if (tile != 0) {
                // common case: none of the lane's 4 candidates beats the current k-th best (d <= tau is a
                // superset of the exact (dist, index) test below)
                const float mn = fminf(fminf(dd[0], dd[1]), fminf(dd[2], dd[3]));
                if (__ballot_sync(0xffffffffu, mn <= tau) == 0u) continue;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (tile == 0 && t == 0) continue;                  // already in the sorted seed
                const int j = jb + t;
                const float d = dd[t];
                bool pass = lex_less(d, j, tau, tauj);
                unsigned m = __ballot_sync(0xffffffffu, pass);
                while (m) {
                    const int src = __ffs(m) - 1;
                    const float nd = __shfl_sync(0xffffffffu, d, src);
                    const int nj = __shfl_sync(0xffffffffu, j, src);
                    const bool before = (lane < K1) && lex_less(ld[u], lj[u], nd, nj);
                    const int pos = __popc(__ballot_sync(0xffffffffu, before));
                    const float upd = __shfl_up_sync(0xffffffffu, ld[u], 1);
                    const int upj = __shfl_up_sync(0xffffffffu, lj[u], 1);
                    if (lane == pos) { ld[u] = nd; lj[u] = nj; }
                    else if (lane > pos && lane < K1) { ld[u] = upd; lj[u] = upj; }
                    tau = __shfl_sync(0xffffffffu, ld[u], K1 - 1);
                    tauj = __shfl_sync(0xffffffffu, lj[u], K1 - 1);
                    pass = pass && (lane != src) && lex_less(d, j, tau, tauj);
                    m = __ballot_sync(0xffffffffu, pass);
                }
            }
        }
        // the next iteration's top-of-loop barrier orders these reads of s.d before the next tile's writes
    }

#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int qq = warp * 8 + u;
        const int i = i0 + qq;
        if (lane >= 1 && lane < K1) {
            const int jn = valid_nbr(lj[u], i, N);
            nbr[qq][lane - 1] = jn;
            if (i < N) idx[((int64_t)b * N + i) * k + (lane - 1)] = jn;
        }
    }
    if (ee == nullptr) return;
    __syncthreads();
    const int nq = min(QT, N - i0);
    const int per_ch = nq * k;
    for (int c = 0; c < C; ++c) {
        const float* row = xb + (int64_t)c * N;
        float* e_ctr = ee + (((int64_t)b * 2 * C + c) * N + i0) * k;
        float* e_dif = ee + (((int64_t)b * 2 * C + C + c) * N + i0) * k;
        for (int e = tid; e < per_ch; e += KNN_THREADS) {
            const int qq = e / k, r = e - qq * k;
            const float ctr = __ldg(row + i0 + qq);
            const float nb = __ldg(row + nbr[qq][r]);
            e_ctr[e] = ctr;
            e_dif[e] = nb - ctr;
        }
    }
}

__global__ void group_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int B, int C, int N,
                             int k, float* __restrict__ ee) {
    // grid: (ceil(N*k / 256), C, B)
    const int b = blockIdx.z, c = blockIdx.y;
    const int64_t per = (int64_t)N * k;
    const float* row = x + ((int64_t)b * C + c) * N;
    const int32_t* ib = idx + (int64_t)b * per;
    float* e_ctr = ee + ((int64_t)b * 2 * C + c) * per;
    float* e_dif = ee + ((int64_t)b * 2 * C + C + c) * per;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < per; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / k);
        const float ctr = __ldg(row + i);
        const float nb = __ldg(row + ib[e]);
        e_ctr[e] = ctr;
        e_dif[e] = nb - ctr;
    }
}

template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (D)src[i];
}

}  // namespace

// the same reduction orders for point-major rows [B*N, C] (channel stride 1)
__global__ void sqnorm_pm_kernel(const float* __restrict__ rows, int B, int C, int N, int main_cols, float* __restrict__ xs) {
    const int64_t total = (int64_t)B * N;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(t % N);
        const float* col = rows + t * C;
        xs[t] = (n < main_cols) ? sqnorm_cascade(col, C, 1) : sqnorm_ilp4(col, C, 1);
    }
}
extern "C" int spgan_sqnorm_pm(const float* rows, int B, int C, int N, int main_cols, float* xs, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(rows && xs && B >= 1 && C >= 1 && N >= 1);
    if (main_cols < 0) main_cols = (N / 32) * 32;
    const int64_t total = (int64_t)B * N;
    sqnorm_pm_kernel<<<ew_grid(total, 128), 128, 0, as_stream(stream)>>>(rows, B, C, N, main_cols, xs);
    return spgan_launch_status();
}
extern "C" int spgan_sqnorm(const float* x, int B, int C, int N, int main_cols, float* xs, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(x && xs && B >= 0 && C >= 1 && N >= 1);
    if (B == 0) return SPGAN_OK;
    if (main_cols < 0) main_cols = (N / 32) * 32;
    const int64_t total = (int64_t)B * N;
    sqnorm_kernel<<<ew_grid(total, 128), 128, 0, as_stream(stream)>>>(x, B, C, N, main_cols, xs);
    return spgan_launch_status();
}

extern "C" int spgan_knn_group(const float* x, const float* xs, int B, int C, int N, int k, int32_t* idx,
                               float* ee, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(x && xs && idx && B >= 0 && C >= 1 && N >= 1 && k >= 1);
    if (k + 1 > N) return SPGAN_E_BADARG;
    if (k + 1 > 32) return SPGAN_E_UNSUPPORTED;
    if (B == 0) return SPGAN_OK;
    static_assert(sizeof(KnnSmem) <= 100 * 1024, "two CTAs per SM");
    cudaError_t e = cudaFuncSetAttribute(knn_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(KnnSmem));
    if (e != cudaSuccess) return (int)e;
    const int q_tiles = (N + QT - 1) / QT;
    const int64_t grid = (int64_t)B * q_tiles;
    if (grid > 0x7fffffffLL) return SPGAN_E_UNSUPPORTED;
    if (N % 4 == 0 && N >= CT && C <= QMAXC && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        static_assert(sizeof(KnnFastSmem) <= 104 * 1024, "two CTAs per SM");
        e = cudaFuncSetAttribute(knn_group_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(KnnFastSmem));
        if (e != cudaSuccess) return (int)e;
        knn_group_fast_kernel<<<(unsigned)grid, KNN_THREADS, sizeof(KnnFastSmem), as_stream(stream)>>>(x, xs, B, C, N, k,
                                                                                                     idx, ee);
        return spgan_launch_status();
    }
    knn_group_kernel<<<(unsigned)grid, KNN_THREADS, sizeof(KnnSmem), as_stream(stream)>>>(x, xs, B, C, N, k, idx, ee);
    return spgan_launch_status();
}

extern "C" int spgan_group(const float* x, const int32_t* idx, int B, int C, int N, int k, float* ee,
                           spgan_stream_t stream) {
    SPGAN_CHECK_ARG(x && idx && ee && B >= 0 && C >= 1 && N >= 1 && k >= 1);
    if (B == 0) return SPGAN_OK;
    if (C > 65535 || B > 65535) return SPGAN_E_UNSUPPORTED;
    const int64_t per = (int64_t)N * k;
    dim3 grid((unsigned)((per + 255) / 256 > 1024 ? 1024 : (per + 255) / 256), C, B);
    group_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, idx, B, C, N, k, ee);
    return spgan_launch_status();
}

extern "C" int spgan_idx32_to_idx64(const int32_t* src, int64_t* dst, int64_t n, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(src && dst && n >= 0);
    if (n == 0) return SPGAN_OK;
    convert_kernel<int32_t, int64_t><<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(src, dst, n);
    return spgan_launch_status();
}

extern "C" int spgan_idx64_to_idx32(const int64_t* src, int32_t* dst, int64_t n, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(src && dst && n >= 0);
    if (n == 0) return SPGAN_OK;
    convert_kernel<int64_t, int32_t><<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(src, dst, n);
    return spgan_launch_status();
}
