// The other kNN / grouping entry points of the reference that share the hot path's arithmetic:
//   knn / get_graph_feature              Generation/modules.py:640-680, Common/ops.py:129-162
//   pairwise_dist                        Generation/modules.py:629-637
//   square_distance / index_points / knn_point   Common/pointnet_util.py:19-60, Common/pointconv_util.py:107-118
//   get_edge_features_xyz                Generation/modules.py:727-776
// They differ from get_edge_features (knn.cu) in three conventions only: the query and candidate sets may be
// different clouds, rank 0 (normally the point itself) may be kept, and the two squared norms may be added in
// the opposite order.  Same tiling as knn_group_kernel: 64 queries x 128 candidates per tile, fp32 FMA chains
// in channel order (bit-exact against the CPU reference), (dist, index)-sorted lists spread over warp lanes.
#include "common.cuh"
#include <float.h>

namespace {

constexpr int QT = 64;       // queries per CTA
constexpr int CT = 128;      // candidates per tile
constexpr int CK = 32;       // channels per staged chunk
constexpr int THREADS = 256;
constexpr int DPAD = 4;

struct QuerySmem {
    float q[CK][QT];
    float c[CK][CT];
    float d[QT][CT + DPAD];
    float xs_q[QT];
    float xs_c[CT];
};

__device__ __forceinline__ bool lex_less(float d0, int j0, float d1, int j1) {
    return d0 < d1 || (d0 == d1 && j0 < j1);
}

// WRITE_DIST: materialise the [Nq, Nc] squared-distance tile instead of selecting neighbours.
template <bool WRITE_DIST>
__global__ void __launch_bounds__(THREADS, 2)
knn_query_kernel(const float* __restrict__ xq, const float* __restrict__ xsq, int Nq,
                 const float* __restrict__ xc, const float* __restrict__ xsc, int Nc, int C, int k, int first_rank,
                 int cand_norm_first, int32_t* __restrict__ idx, float* __restrict__ dist_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QuerySmem& s = *reinterpret_cast<QuerySmem*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;     // ty -> 4 queries, tx -> 8 candidates
    const int q_tiles = (Nq + QT - 1) / QT;
    const int b = blockIdx.x / q_tiles;
    const int i0 = (blockIdx.x % q_tiles) * QT;
    const float* xqb = xq + (int64_t)b * C * Nq;
    const float* xcb = xc + (int64_t)b * C * Nc;
    const int K1 = k + first_rank;              // list length: ranks 0 .. K1-1

    if (tid < QT) s.xs_q[tid] = (i0 + tid < Nq) ? xsq[(int64_t)b * Nq + i0 + tid] : 0.f;

    float ld[8];
    int lj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { ld[u] = FLT_MAX; lj[u] = 0x7fffffff; }

    for (int j0 = 0; j0 < Nc; j0 += CT) {
        float acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;

        for (int c0 = 0; c0 < C; c0 += CK) {
            __syncthreads();
            for (int e = tid; e < CK * QT; e += THREADS) {
                const int cc = e / QT, qq = e % QT;
                const int ch = c0 + cc, i = i0 + qq;
                s.q[cc][qq] = (ch < C && i < Nq) ? __ldg(xqb + (int64_t)ch * Nq + i) : 0.f;
            }
            for (int e = tid; e < CK * CT; e += THREADS) {
                const int cc = e / CT, jj = e % CT;
                const int ch = c0 + cc, j = j0 + jj;
                s.c[cc][jj] = (ch < C && j < Nc) ? __ldg(xcb + (int64_t)ch * Nc + j) : 0.f;
            }
            if (c0 == 0 && tid < CT) s.xs_c[tid] = (j0 + tid < Nc) ? xsc[(int64_t)b * Nc + j0 + tid] : 0.f;
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
        // dist = (-2*dot + first) + second, two separate roundings; which norm goes first is the caller's
        // convention (modules.py:699 / pointnet_util.py:37-39: query first; modules.py:643 `knn`: candidate first)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float xq_n = s.xs_q[ty * 4 + a];
            float out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int jj = (c < 4) ? tx * 4 + c : 64 + tx * 4 + (c - 4);
                const float xc_n = s.xs_c[jj];
                const float m2 = __fmul_rn(-2.0f, acc[a][c]);
                const float dq = cand_norm_first ? __fadd_rn(__fadd_rn(m2, xc_n), xq_n) : __fadd_rn(__fadd_rn(m2, xq_n), xc_n);
                out[c] = WRITE_DIST ? dq : fminf(dq, FLT_MAX);      // selection key: NaN/+inf ordered by index (knn.cu ord_key)
            }
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][tx * 4]) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(&s.d[ty * 4 + a][64 + tx * 4]) = make_float4(out[4], out[5], out[6], out[7]);
        }
        __syncthreads();

        if (WRITE_DIST) {
            // rows of the tile -> global, 128 consecutive floats per query (coalesced)
            const int nc = min(CT, Nc - j0);
            for (int e = tid; e < QT * CT; e += THREADS) {
                const int qq = e / CT, jj = e % CT;
                if (i0 + qq < Nq && jj < nc)
                    dist_out[((int64_t)b * Nq + i0 + qq) * Nc + j0 + jj] = s.d[qq][jj];
            }
            continue;       // the next tile's first barrier orders these reads before its writes of s.d
        }

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
                bool pass = (j < Nc) && lex_less(d, j, tau, tauj);
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
    }
    if (WRITE_DIST) return;

#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int i = i0 + warp * 8 + u;
        if (lane >= first_rank && lane < K1 && i < Nq)
            idx[((int64_t)b * Nq + i) * k + (lane - first_rank)] = lj[u] < Nc ? lj[u] : min((int)i, Nc - 1);
    }
}

// |x_r|^2 of point-major rows, rounded squares added in channel order (pointnet_util.py:38-39 on xyz rows).
__global__ void sqnorm_rows_kernel(const float* __restrict__ x, int64_t R, int C, float* __restrict__ xs) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
        const float* p = x + r * C;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(p + c);
            acc = __fadd_rn(acc, __fmul_rn(v, v));
        }
        xs[r] = acc;
    }
}

// ee[b, :, n, r]: the two channel halves of the grouped tensor in either order.
__global__ void group_ex_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int C, int N, int k,
                                int diff_first, float* __restrict__ ee) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int64_t per = (int64_t)N * k;
    const float* row = x + ((int64_t)b * C + c) * N;
    const int32_t* ib = idx + (int64_t)b * per;
    float* e_ctr = ee + ((int64_t)b * 2 * C + (diff_first ? C : 0) + c) * per;
    float* e_dif = ee + ((int64_t)b * 2 * C + (diff_first ? 0 : C) + c) * per;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < per; e += (int64_t)gridDim.x * blockDim.x) {
        const float ctr = __ldg(row + e / k);
        const float nb = __ldg(row + ib[e]);
        e_ctr[e] = ctr;
        e_dif[e] = nb - ctr;
    }
}

// out[b, s, :] = points[b, idx[b, s], :]; one thread per (row, 4-channel group) when C % 4 == 0.
template <typename I, int V>
__global__ void gather_rows_kernel(const float* __restrict__ points, const I* __restrict__ idx, int N, int64_t S,
                                   int C, int64_t total_rows, float* __restrict__ out, int* __restrict__ status) {
    const int cv = C / V;
    const int64_t total = total_rows * cv;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / cv;
        const int c = (int)(t - r * cv) * V;
        const int64_t b = r / S;
        int64_t j = (int64_t)idx[r];
        if (j < 0) j += N;          // negative indices wrap like the reference's fancy indexing
        if (j < 0 || j >= N) { if (status) atomicExch(status, 1); continue; }
        const float* src = points + (b * N + j) * C + c;
        float* dst = out + r * C + c;
        if (V == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
        else *dst = __ldg(src);
    }
}

// dpoints[b, idx[b, s], :] += g[b, s, :]
template <typename I>
__global__ void scatter_add_rows_kernel(const float* __restrict__ g, const I* __restrict__ idx, int N, int64_t S,
                                        int C, int64_t total_rows, float* __restrict__ dpoints) {
    const int64_t total = total_rows * C;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / C;
        const int c = (int)(t - r * C);
        const int64_t b = r / S;
        int64_t j = (int64_t)idx[r];
        if (j < 0) j += N;
        if (j < 0 || j >= N) continue;
        atomicAdd(dpoints + (b * N + j) * C + c, g[t]);
    }
}

int launch_query(bool write_dist, const float* xq, const float* xsq, int Nq, const float* xc, const float* xsc, int Nc,
                 int B, int C, int k, int first_rank, int cand_norm_first, int32_t* idx, float* dist,
                 cudaStream_t st) {
    static_assert(sizeof(QuerySmem) <= 100 * 1024, "two CTAs per SM");
    const int64_t grid = (int64_t)B * ((Nq + QT - 1) / QT);
    if (grid > 0x7fffffffLL) return SPGAN_E_UNSUPPORTED;
    cudaError_t e;
    if (write_dist) {
        e = cudaFuncSetAttribute(knn_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuerySmem));
        if (e != cudaSuccess) return (int)e;
        knn_query_kernel<true><<<(unsigned)grid, THREADS, sizeof(QuerySmem), st>>>(xq, xsq, Nq, xc, xsc, Nc, C, 1, 0,
                                                                                  cand_norm_first, nullptr, dist);
    } else {
        e = cudaFuncSetAttribute(knn_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuerySmem));
        if (e != cudaSuccess) return (int)e;
        knn_query_kernel<false><<<(unsigned)grid, THREADS, sizeof(QuerySmem), st>>>(xq, xsq, Nq, xc, xsc, Nc, C, k,
                                                                                   first_rank, cand_norm_first, idx,
                                                                                   nullptr);
    }
    return spgan_launch_status();
}

}  // namespace

extern "C" int spgan_sqnorm_rows(const float* x_rows, int64_t R, int C, float* xs, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(x_rows && xs && R >= 0 && C >= 1);
    if (R == 0) return SPGAN_OK;
    sqnorm_rows_kernel<<<ew_grid(R, 256), 256, 0, as_stream(stream)>>>(x_rows, R, C, xs);
    return spgan_launch_status();
}

extern "C" int spgan_knn_query(const float* xq_bcn, const float* xsq, int Nq, const float* xc_bcn, const float* xsc,
                               int Nc, int B, int C, int k, int first_rank, int cand_norm_first, int32_t* idx,
                               spgan_stream_t stream) {
    SPGAN_CHECK_ARG(xq_bcn && xsq && xc_bcn && xsc && idx && B >= 0 && C >= 1 && Nq >= 0 && Nc >= 1 && k >= 1);
    SPGAN_CHECK_ARG(first_rank == 0 || first_rank == 1);
    if (k + first_rank > Nc) return SPGAN_E_BADARG;
    if (k + first_rank > 32) return SPGAN_E_UNSUPPORTED;
    if (B == 0 || Nq == 0) return SPGAN_OK;
    return launch_query(false, xq_bcn, xsq, Nq, xc_bcn, xsc, Nc, B, C, k, first_rank, cand_norm_first ? 1 : 0, idx,
                        nullptr, as_stream(stream));
}

extern "C" int spgan_pairwise_sqdist(const float* xq_bcn, const float* xsq, int Nq, const float* xc_bcn,
                                     const float* xsc, int Nc, int B, int C, int cand_norm_first, float* dist,
                                     spgan_stream_t stream) {
    SPGAN_CHECK_ARG(xq_bcn && xsq && xc_bcn && xsc && dist && B >= 0 && C >= 1 && Nq >= 0 && Nc >= 0);
    if (B == 0 || Nq == 0 || Nc == 0) return SPGAN_OK;
    return launch_query(true, xq_bcn, xsq, Nq, xc_bcn, xsc, Nc, B, C, 1, 0, cand_norm_first ? 1 : 0, nullptr, dist,
                        as_stream(stream));
}

extern "C" int spgan_group_ex(const float* x_bcn, const int32_t* idx, int B, int C, int N, int k, int diff_first,
                              float* ee, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(x_bcn && idx && ee && B >= 0 && C >= 1 && N >= 1 && k >= 1);
    if (B == 0) return SPGAN_OK;
    if (C > 65535 || B > 65535) return SPGAN_E_UNSUPPORTED;
    const int64_t per = (int64_t)N * k;
    dim3 grid((unsigned)((per + 255) / 256 > 1024 ? 1024 : (per + 255) / 256), C, B);
    group_ex_kernel<<<grid, 256, 0, as_stream(stream)>>>(x_bcn, idx, C, N, k, diff_first ? 1 : 0, ee);
    return spgan_launch_status();
}

extern "C" int spgan_gather_rows(const float* points, const void* idx, int idx_is_int64, int B, int N, int64_t S, int C,
                                 float* out, int* status, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(points && idx && out && B >= 0 && N >= 1 && S >= 0 && C >= 1);
    const int64_t rows = (int64_t)B * S;
    if (rows == 0) return SPGAN_OK;
    cudaStream_t st = as_stream(stream);
    const bool v4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int g = ew_grid(rows * (v4 ? C / 4 : C), 256);
    if (idx_is_int64) {
        if (v4) gather_rows_kernel<int64_t, 4><<<g, 256, 0, st>>>(points, (const int64_t*)idx, N, S, C, rows, out, status);
        else gather_rows_kernel<int64_t, 1><<<g, 256, 0, st>>>(points, (const int64_t*)idx, N, S, C, rows, out, status);
    } else {
        if (v4) gather_rows_kernel<int32_t, 4><<<g, 256, 0, st>>>(points, (const int32_t*)idx, N, S, C, rows, out, status);
        else gather_rows_kernel<int32_t, 1><<<g, 256, 0, st>>>(points, (const int32_t*)idx, N, S, C, rows, out, status);
    }
    return spgan_launch_status();
}

extern "C" int spgan_scatter_add_rows(const float* g, const void* idx, int idx_is_int64, int B, int N, int64_t S, int C,
                                      float* dpoints, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(g && idx && dpoints && B >= 0 && N >= 1 && S >= 0 && C >= 1);
    const int64_t rows = (int64_t)B * S;
    if (rows == 0) return SPGAN_OK;
    cudaStream_t st = as_stream(stream);
    const int grid = ew_grid(rows * C, 256);
    if (idx_is_int64) scatter_add_rows_kernel<int64_t><<<grid, 256, 0, st>>>(g, (const int64_t*)idx, N, S, C, rows, dpoints);
    else scatter_add_rows_kernel<int32_t><<<grid, 256, 0, st>>>(g, (const int32_t*)idx, N, S, C, rows, dpoints);
    return spgan_launch_status();
}
