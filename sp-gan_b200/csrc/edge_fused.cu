// EdgeBlock passes with the train-mode BatchNorm reductions folded into the kernels that already hold the values
// (Generation/Generator.py:75-88).  Each kernel below removes full [E, F] tensor passes of the unfused chain:
//
//   * edge_combine_stats   : the edge gather w0 = p1[nbr] - p1[p] + b (or y = a[p] + d[nbr] - d[p] + b) ALSO leaves the
//                            per-column sum / sum of squares of what it writes as deterministic per-CTA partial rows
//                            (spgan_bn_finalize turns them into the batch statistics): the statistics pass over the
//                            [E, F] tensor disappears.
//   * attn_bwd_stats       : backward of prod = lrelu(bn(xy)) * softmax_k(lrelu(bn(xw))) up to the two activated
//                            tensors, plus the four column sums (sum g', sum g' xhat for both BatchNorms) the BatchNorm
//                            backward needs -- the two norm_bwd_reduce passes (2 tensor reads each) disappear for one
//                            extra read of xw.
//   * edge_combine_bwd_bn  : the BatchNorm + LeakyReLU backward of the conv_x branch applied on the fly inside the
//                            scatter of the edge gather's backward: d(xy) is never written.
//   * partials_finalize    : fixed-tree fp64 reduction of per-CTA partial rows (deterministic), optional in-place
//                            accumulation into parameter gradients.
#include "common.cuh"
#include <float.h>

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float bn_act_pre(float x, float m, float r, float g, float b) { return fmaf((x - m) * r, g, b); }

// one thread per (point, 4-channel group); 256 % (C / 4) == 0, so a thread keeps its channel group over the
// grid-stride loop and the block-level fold is a fixed-order sum over the threads of equal channel group
template <bool STORE>
__global__ void __launch_bounds__(256)
edge_combine_stats_kernel(const float* __restrict__ pc, const float* __restrict__ pn, const int32_t* __restrict__ idx,
                          const float* __restrict__ bias, int64_t P, int N, int k, int C, float* __restrict__ out,
                          float* __restrict__ col_sum, float* __restrict__ col_sqsum) {
    __shared__ float4 red[2][256];
    const int Cw = C >> 2;
    const int64_t total = P * Cw;
    const int cw = threadIdx.x % Cw;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = ldg4(bias + 4 * cw);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / Cw;
        const int64_t base = (p / N) * N;
        const int32_t* ip = idx + p * k;
        // same operation order as spgan_edge_combine: (pn[j] - pn[p]) + pc[p] + bias
        const float4 b = ldg4(pn + p * C + 4 * cw);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pc) q = ldg4(pc + p * C + 4 * cw);
        float4* o = reinterpret_cast<float4*>(out) + p * k * Cw + cw;
#pragma unroll 5
        for (int r = 0; r < k; ++r) {
            const int64_t j = base + __ldg(ip + r);
            const float4 a = ldg4(pn + j * C + 4 * cw);
            float4 v = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
            if (pc) { v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
            if (bias) { v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
            if (STORE) o[(int64_t)r * Cw] = v;
            s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
            s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
        }
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < Cw) {
        float4 a = red[0][threadIdx.x], b = red[1][threadIdx.x];
        for (int t = threadIdx.x + Cw; t < 256; t += Cw) {
            const float4 u = red[0][t], v = red[1][t];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
        reinterpret_cast<float4*>(col_sum + (int64_t)blockIdx.x * C)[threadIdx.x] = a;
        reinterpret_cast<float4*>(col_sqsum + (int64_t)blockIdx.x * C)[threadIdx.x] = b;
    }
}

struct BnColP { const float* mean; const float* rstd; const float* gamma; const float* beta; };

// one thread per (point, channel); 256 % C == 0 (C = 32 .. 256)
template <int KM>
__global__ void __launch_bounds__(256)
attn_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ xw, const float* __restrict__ xy,
                      const float* __restrict__ w, int64_t P, int k, int C, BnColP bw, BnColP by, float slope,
                      float* __restrict__ dwa, float* __restrict__ dya, float* __restrict__ part) {
    __shared__ float red[4][256];
    const int64_t total = P * C;
    const int c = threadIdx.x % C;
    const float mw = __ldg(bw.mean + c), rw = __ldg(bw.rstd + c), gw = __ldg(bw.gamma + c), bbw = __ldg(bw.beta + c);
    const float my = __ldg(by.mean + c), ry = __ldg(by.rstd + c), gy = __ldg(by.gamma + c), bby = __ldg(by.beta + c);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;       // sum g'_w, sum g'_w xhat_w, sum g'_y, sum g'_y xhat_y
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int64_t o = p * k * C + c;
        float gv[KM], wv[KM], yv[KM], xv[KM];
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const int64_t a = o + (int64_t)r * C;
                gv[r] = __ldg(g + a);
                wv[r] = __ldg(w + a);
                yv[r] = __ldg(xy + a);
                xv[r] = __ldg(xw + a);
            }
        float s = 0.f;
        float ya[KM];
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const float pre = bn_act_pre(yv[r], my, ry, gy, bby);
                ya[r] = gv[r] * lrelu_f(pre, slope);                    // g * lrelu(bn(xy))
                s = fmaf(ya[r], wv[r], s);
                // d/d lrelu(bn(xy)) = g w; through the activation: the BatchNorm backward's g'
                const float d = gv[r] * wv[r];
                dya[o + (int64_t)r * C] = d;
                const float gp = pre > 0.f ? d : d * slope;
                a2 += gp;
                a3 = fmaf(gp, (yv[r] - my) * ry, a3);
            }
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const float d = wv[r] * (ya[r] - s);
                dwa[o + (int64_t)r * C] = d;
                const float xh = (xv[r] - mw) * rw;
                const float gp = fmaf(xh, gw, bbw) > 0.f ? d : d * slope;
                a0 += gp;
                a1 = fmaf(gp, xh, a1);
            }
    }
    red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2; red[3][threadIdx.x] = a3;
    __syncthreads();
    for (int t = threadIdx.x; t < 4 * C; t += 256) {
        const int v = t / C, cc = t - v * C;
        float acc = red[v][cc];
        for (int u = cc + C; u < 256; u += C) acc += red[v][u];
        part[((int64_t)blockIdx.x * 4 + v) * C + cc] = acc;
    }
}

// part [rows, nvals, C] -> out [nvals, C] (double, fixed tree: 32 row lanes in order, then the 32 lane sums in order)
__global__ void __launch_bounds__(1024)
partials_finalize_kernel(const float* __restrict__ part, int rows, int nvals, int C, float* __restrict__ out,
                         float* acc0, float* acc1, float* acc2, float* acc3) {
    __shared__ double red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int v = blockIdx.y;
    double s = 0.0;
    if (c < C)
        for (int r = ty; r < rows; r += 32) s += (double)__ldg(part + ((int64_t)r * nvals + v) * C + c);
    red[ty][tx] = s;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    s = 0.0;
    for (int t = 0; t < 32; ++t) s += red[t][tx];
    out[v * C + c] = (float)s;
    float* acc = v == 0 ? acc0 : v == 1 ? acc1 : v == 2 ? acc2 : acc3;
    if (acc) acc[c] += (float)s;
}

// dx = gamma rstd (g' - sg/n - xhat sgx/n) formed per element from (g, x) and scattered like spgan_edge_combine_bwd
__global__ void __launch_bounds__(256)
edge_combine_bwd_bn_kernel(const float* __restrict__ g, const float* __restrict__ x, const int32_t* __restrict__ idx,
                           int64_t P, int N, int k, int C, BnColP bn, const float* __restrict__ sg,
                           const float* __restrict__ sgx, float inv_n, float slope, float* __restrict__ dpc,
                           float* __restrict__ dpn) {
    const int Cw = C >> 2;
    const int64_t total = P * Cw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cw = (int)(i % Cw);
        const int64_t p = i / Cw;
        const int64_t base = (p / N) * N;
        const float4 m = ldg4(bn.mean + 4 * cw), rs = ldg4(bn.rstd + 4 * cw), gm = ldg4(bn.gamma + 4 * cw),
                     bt = ldg4(bn.beta + 4 * cw);
        float4 a = ldg4(sg + 4 * cw), b = ldg4(sgx + 4 * cw);
        a.x *= inv_n; a.y *= inv_n; a.z *= inv_n; a.w *= inv_n;
        b.x *= inv_n; b.y *= inv_n; b.z *= inv_n; b.w *= inv_n;
        const float4 coef = make_float4(gm.x * rs.x, gm.y * rs.y, gm.z * rs.z, gm.w * rs.w);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
        for (int r = 0; r < k; ++r) {
            const int64_t e = (p * k + r) * C + 4 * cw;
            float4 gi = ldg4(g + e);
            const float4 xv = ldg4(x + e);
            const float4 xh = make_float4((xv.x - m.x) * rs.x, (xv.y - m.y) * rs.y, (xv.z - m.z) * rs.z, (xv.w - m.w) * rs.w);
            // same operation order as NormBwdApplyOp4: mask, then coef * ((g' - a) - xh * b)
            if (!(fmaf(xh.x, gm.x, bt.x) > 0.f)) gi.x *= slope;
            if (!(fmaf(xh.y, gm.y, bt.y) > 0.f)) gi.y *= slope;
            if (!(fmaf(xh.z, gm.z, bt.z) > 0.f)) gi.z *= slope;
            if (!(fmaf(xh.w, gm.w, bt.w) > 0.f)) gi.w *= slope;
            float4 v;
            v.x = coef.x * ((gi.x - a.x) - xh.x * b.x);
            v.y = coef.y * ((gi.y - a.y) - xh.y * b.y);
            v.z = coef.z * ((gi.z - a.z) - xh.z * b.z);
            v.w = coef.w * ((gi.w - a.w) - xh.w * b.w);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            const int64_t j = base + __ldg(idx + p * k + r);
            atomicAdd(reinterpret_cast<float4*>(dpn + j * C) + cw, v);
        }
        if (dpc) reinterpret_cast<float4*>(dpc + p * C)[cw] = s;
        atomicAdd(reinterpret_cast<float4*>(dpn + p * C) + cw, make_float4(-s.x, -s.y, -s.z, -s.w));
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" size_t spgan_edge_stats_rows(int64_t P, int C) {
    if (P <= 0 || C < 4 || (C & 3) != 0 || 256 % (C >> 2) != 0) return 0;
    // 4 CTAs per SM are resident (60 registers): one grid-stride wave.  More CTAs would only mean more partial rows for
    // the finalize kernel to read (it was 2368 rows = 18 us per BatchNorm; 592 rows = 5 us)
    return (size_t)ew_grid(P * (C >> 2), 256, 4);
}

extern "C" int spgan_edge_combine_stats(const float* pc, const float* pn, const int32_t* idx, const float* bias, int64_t P,
                                        int N, int k, int C, float* out, float* col_sum, float* col_sqsum,
                                        spgan_stream_t s) {
    SPGAN_CHECK_ARG(pn && idx && col_sum && col_sqsum && P >= 1 && N >= 1 && k >= 1 && C >= 1 && P % N == 0);
    const int rows = (int)spgan_edge_stats_rows(P, C);
    if (rows == 0 || !al16(pn) || (pc && !al16(pc)) || (bias && !al16(bias)) || (out && !al16(out)) || !al16(col_sum) ||
        !al16(col_sqsum))
        return SPGAN_E_UNSUPPORTED;
    if (out)
        edge_combine_stats_kernel<true><<<rows, 256, 0, as_stream(s)>>>(pc, pn, idx, bias, P, N, k, C, out, col_sum, col_sqsum);
    else
        edge_combine_stats_kernel<false><<<rows, 256, 0, as_stream(s)>>>(pc, pn, idx, bias, P, N, k, C, out, col_sum, col_sqsum);
    return spgan_launch_status();
}

extern "C" size_t spgan_attn_bwd_rows(int64_t P, int k, int C) {
    if (P <= 0 || k < 1 || k > 16 || C < 1 || C > 256 || 256 % C != 0) return 0;
    return (size_t)ew_grid(P * C, 256, 2);          // 2 CTAs per SM are resident (90 registers): one grid-stride wave
}

extern "C" int spgan_bn_softmax_mul_k_bwd_stats(const float* g, const float* xw, const float* xy, const float* w, int64_t P,
                                                int k, int C, const float* mean_w, const float* rstd_w,
                                                const float* gamma_w, const float* beta_w, const float* mean_y,
                                                const float* rstd_y, const float* gamma_y, const float* beta_y,
                                                float slope, float* dwa, float* dya, float* part, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && xw && xy && w && dwa && dya && part && mean_w && rstd_w && gamma_w && beta_w && mean_y && rstd_y &&
                    gamma_y && beta_y);
    const int rows = (int)spgan_attn_bwd_rows(P, k, C);
    if (rows == 0) return SPGAN_E_UNSUPPORTED;
    const BnColP bw{mean_w, rstd_w, gamma_w, beta_w}, by{mean_y, rstd_y, gamma_y, beta_y};
    if (k <= 10)
        attn_bwd_stats_kernel<10><<<rows, 256, 0, as_stream(s)>>>(g, xw, xy, w, P, k, C, bw, by, slope, dwa, dya, part);
    else
        attn_bwd_stats_kernel<16><<<rows, 256, 0, as_stream(s)>>>(g, xw, xy, w, P, k, C, bw, by, slope, dwa, dya, part);
    return spgan_launch_status();
}

extern "C" int spgan_partials_finalize(const float* part, int rows, int nvals, int C, float* out, float* acc0, float* acc1,
                                       float* acc2, float* acc3, spgan_stream_t s) {
    SPGAN_CHECK_ARG(part && out && rows >= 1 && nvals >= 1 && nvals <= 4 && C >= 1);
    partials_finalize_kernel<<<dim3((C + 31) / 32, nvals), 1024, 0, as_stream(s)>>>(part, rows, nvals, C, out, acc0, acc1,
                                                                                    acc2, acc3);
    return spgan_launch_status();
}

extern "C" int spgan_edge_combine_bwd_bn(const float* g, const float* x, const int32_t* idx, int64_t P, int N, int k, int C,
                                         const float* mean, const float* rstd, const float* gamma, const float* beta,
                                         const float* sg, const float* sgx, float slope, float* dpc, float* dpn,
                                         spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && idx && dpn && mean && rstd && gamma && beta && sg && sgx && P >= 1 && N >= 1 && k >= 1 &&
                    P % N == 0);
    if ((C & 3) != 0 || !al16(g) || !al16(x) || !al16(dpn) || (dpc && !al16(dpc)) || !al16(mean) || !al16(rstd) ||
        !al16(gamma) || !al16(beta) || !al16(sg) || !al16(sgx))
        return SPGAN_E_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(dpn, 0, (size_t)P * C * sizeof(float), as_stream(s));
    if (e != cudaSuccess) return (int)e;
    const float inv_n = 1.f / (float)((double)P * k);
    const BnColP bn{mean, rstd, gamma, beta};
    edge_combine_bwd_bn_kernel<<<ew_grid(P * (C / 4), 256, 16), 256, 0, as_stream(s)>>>(g, x, idx, P, N, k, C, bn, sg, sgx,
                                                                                       inv_n, slope, dpc, dpn);
    return spgan_launch_status();
}
