// Fused train-mode BatchNorm + LeakyReLU + max over the points of each cloud, for the critic's last
// point-wise layer (Generation/Discriminator.py:77-81,104: Conv1d -> BatchNorm1d -> LeakyReLU ->
// adaptive_max_pool1d).  The [B*N, C] activation is read ONCE in the forward (column moments and the
// per-cloud max / min with their first positions in the same pass) and the normalised tensor is never
// materialised: every fp32 operation of y = lrelu(fma((x - mean) * rstd, gamma, beta)) is monotone in x, so
// max_n y = y(max_n x) for gamma >= 0 and y(min_n x) for gamma < 0.  The backward is sparse in the incoming
// gradient (one row per (cloud, channel)): the two BatchNorm sums come from B*C gathered values and the
// dense pass reads x and writes dx only.
//
// Tie rule: the pooled VALUE equals the reference's; the arg max is the first row attaining the extreme
// of x, which can differ from the first row attaining the extreme of y only when two different x round to
// the same y.
#include "common.cuh"
#include "norm_fast.cuh"
#include <float.h>

namespace {
using namespace fastnorm;

constexpr int PTX = 32;                  // float4 column groups per block (128 columns)
constexpr int PTY = 8;                   // row lanes per block
constexpr int NOIDX = 0x7fffffff;

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ void upd_max(float v, int r, float& best, int& bi) {
    if (v > best || (v == best && r < bi)) { best = v; bi = r; }
}
__device__ __forceinline__ void upd_min(float v, int r, float& best, int& bi) {
    if (v < best || (v == best && r < bi)) { best = v; bi = r; }
}

// partial layout, by = seg * chunks + chunk:
//   psum [by][2][C]  shifted moments (d = x - x[row 0]),  pmax/pmin [by][C],  pamax/pamin [by][C] (row in segment)
__global__ void __launch_bounds__(PTX * PTY)
bn_pool_partial_kernel(const float* __restrict__ x, int C4, int64_t seg_rows, int chunks, int64_t rows_per_chunk,
                       float* __restrict__ psum, float* __restrict__ pmax, int* __restrict__ pamax,
                       float* __restrict__ pmin, int* __restrict__ pamin) {
    __shared__ float4 rs[PTY][2][PTX];
    __shared__ float4 rmx[PTY][PTX], rmn[PTY][PTX];
    __shared__ int4 ramx[PTY][PTX], ramn[PTY][PTX];
    const int tx = threadIdx.x % PTX, ty = threadIdx.x / PTX;
    const int c4 = blockIdx.x * PTX + tx;
    const int C = C4 * 4;
    const int64_t seg = blockIdx.y / chunks;
    const int chunk = blockIdx.y % chunks;
    const int64_t r0 = (int64_t)chunk * rows_per_chunk;
    int64_t r1 = r0 + rows_per_chunk;
    if (r1 > seg_rows) r1 = seg_rows;
    float4 s0 = f4(0.f), s1 = f4(0.f), mx = f4(-INFINITY), mn = f4(INFINITY);
    int4 amx = make_int4(NOIDX, NOIDX, NOIDX, NOIDX), amn = amx;
    if (c4 < C4) {
        const float4 shift = ld4(x + c4 * 4);
        const float* base = x + seg * seg_rows * C + c4 * 4;
#pragma unroll 4
        for (int64_t r = r0 + ty; r < r1; r += PTY) {
            const float4 v = ld4(base + r * C);
            const float4 d = sub4(v, shift);
            s0 = add4(s0, d);
            s1 = fma4(d, d, s1);
            const int ri = (int)r;
            if (v.x > mx.x) { mx.x = v.x; amx.x = ri; }
            if (v.y > mx.y) { mx.y = v.y; amx.y = ri; }
            if (v.z > mx.z) { mx.z = v.z; amx.z = ri; }
            if (v.w > mx.w) { mx.w = v.w; amx.w = ri; }
            if (v.x < mn.x) { mn.x = v.x; amn.x = ri; }
            if (v.y < mn.y) { mn.y = v.y; amn.y = ri; }
            if (v.z < mn.z) { mn.z = v.z; amn.z = ri; }
            if (v.w < mn.w) { mn.w = v.w; amn.w = ri; }
        }
    }
    rs[ty][0][tx] = s0; rs[ty][1][tx] = s1;
    rmx[ty][tx] = mx; rmn[ty][tx] = mn; ramx[ty][tx] = amx; ramn[ty][tx] = amn;
    __syncthreads();
    if (ty == 0 && c4 < C4) {
#pragma unroll
        for (int t = 1; t < PTY; ++t) {
            s0 = add4(s0, rs[t][0][tx]);
            s1 = add4(s1, rs[t][1][tx]);
            const float4 v = rmx[t][tx]; const int4 j = ramx[t][tx];
            upd_max(v.x, j.x, mx.x, amx.x); upd_max(v.y, j.y, mx.y, amx.y);
            upd_max(v.z, j.z, mx.z, amx.z); upd_max(v.w, j.w, mx.w, amx.w);
            const float4 w = rmn[t][tx]; const int4 i = ramn[t][tx];
            upd_min(w.x, i.x, mn.x, amn.x); upd_min(w.y, i.y, mn.y, amn.y);
            upd_min(w.z, i.z, mn.z, amn.z); upd_min(w.w, i.w, mn.w, amn.w);
        }
        const int64_t by = blockIdx.y;
        st4(psum + (by * 2 + 0) * C + c4 * 4, s0);
        st4(psum + (by * 2 + 1) * C + c4 * 4, s1);
        st4(pmax + by * C + c4 * 4, mx);
        st4(pmin + by * C + c4 * 4, mn);
        *reinterpret_cast<int4*>(pamax + by * C + c4 * 4) = amx;
        *reinterpret_cast<int4*>(pamin + by * C + c4 * 4) = amn;
    }
}

// block = 32 columns x 8 lanes: (1) column moments over all partial rows in double, fixed order;
// (2) per cloud: combine the chunk extremes, pick max or min by the sign of gamma, apply BN + LeakyReLU.
__global__ void __launch_bounds__(256)
bn_pool_final_kernel(const float* __restrict__ x, int C, int64_t nseg, int64_t seg_rows, int chunks, float eps,
                     const float* __restrict__ psum, const float* __restrict__ pmax, const int* __restrict__ pamax,
                     const float* __restrict__ pmin, const int* __restrict__ pamin, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float slope, float* __restrict__ mean, float* __restrict__ rstd,
                     float* __restrict__ var, float* __restrict__ pooled, int32_t* __restrict__ arg) {
    __shared__ double red[8][2][33];
    __shared__ float s_mean[32], s_rstd[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int64_t nby = nseg * chunks;
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
        for (int64_t by = ty; by < nby; by += 8) {
            a0 += (double)__ldg(psum + (by * 2 + 0) * C + c);
            a1 += (double)__ldg(psum + (by * 2 + 1) * C + c);
        }
    }
    red[ty][0][tx] = a0; red[ty][1][tx] = a1;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int t = 1; t < 8; ++t) { a0 += red[t][0][tx]; a1 += red[t][1][tx]; }
        const double n = (double)nseg * (double)seg_rows;
        const double m1 = a0 / n;
        double v = a1 / n - m1 * m1;
        if (v < 0.0) v = 0.0;
        const float m = (float)((double)__ldg(x + c) + m1);
        const float rs = (float)(1.0 / sqrt(v + (double)eps));
        mean[c] = m; rstd[c] = rs;
        if (var) var[c] = (float)v;
        s_mean[tx] = m; s_rstd[tx] = rs;
    }
    __syncthreads();
    if (c >= C) return;
    const float m = s_mean[tx], rs = s_rstd[tx];
    const float gm = gamma ? __ldg(gamma + c) : 1.f, bt = beta ? __ldg(beta + c) : 0.f;
    for (int64_t seg = ty; seg < nseg; seg += 8) {
        float best; int bi;
        if (gm >= 0.f) {
            best = -INFINITY; bi = NOIDX;
            for (int ch = 0; ch < chunks; ++ch)
                upd_max(__ldg(pmax + (seg * chunks + ch) * C + c), __ldg(pamax + (seg * chunks + ch) * C + c), best, bi);
        } else {
            best = INFINITY; bi = NOIDX;
            for (int ch = 0; ch < chunks; ++ch)
                upd_min(__ldg(pmin + (seg * chunks + ch) * C + c), __ldg(pamin + (seg * chunks + ch) * C + c), best, bi);
        }
        if (bi == NOIDX || gm == 0.f) { bi = 0; best = __ldg(x + seg * seg_rows * C + c); }   // constant / NaN column
        const float y = fmaf((best - m) * rs, gm, bt);
        pooled[seg * C + c] = lrelu_f(y, slope);
        arg[seg * C + c] = bi;
    }
}

// gradient w.r.t. the pre-activation at the selected rows, and the two BatchNorm sums over them
__global__ void __launch_bounds__(256)
bn_pool_bwd_sums_kernel(const float* __restrict__ gp, const float* __restrict__ x, const int32_t* __restrict__ arg,
                        int C, int64_t nseg, int64_t seg_rows, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                        float slope, float* __restrict__ gprime, float* __restrict__ sg, float* __restrict__ sgx) {
    __shared__ float red[8][2][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
        const float m = __ldg(mean + c), rs = __ldg(rstd + c);
        const float gm = gamma ? __ldg(gamma + c) : 1.f, bt = beta ? __ldg(beta + c) : 0.f;
        for (int64_t seg = ty; seg < nseg; seg += 8) {
            const int a = __ldg(arg + seg * C + c);
            const float xh = (__ldg(x + (seg * seg_rows + a) * C + c) - m) * rs;
            float g = __ldg(gp + seg * C + c);
            if (!(fmaf(xh, gm, bt) > 0.f)) g *= slope;
            gprime[seg * C + c] = g;
            a0 += g; a1 = fmaf(g, xh, a1);
        }
    }
    red[ty][0][tx] = a0; red[ty][1][tx] = a1;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int t = 1; t < 8; ++t) { a0 += red[t][0][tx]; a1 += red[t][1][tx]; }
        sg[c] = a0; sgx[c] = a1;
    }
}

// dx = gamma * rstd * (g' [only at the selected row] - sum(g')/n - xhat * sum(g' xhat)/n)
struct BnPoolBwdApplyOp4 {
    const float* x; int C; int64_t seg_rows; float inv_n; const float* mean; const float* rstd; const float* gamma;
    const float* sg; const float* sgx; const float* gprime; const int32_t* arg; float* dx;
    struct State { float4 mean, rstd, coef, a, b, g; int4 arg; int64_t base; };
    __device__ State init(int c4, int64_t seg) const {
        const float4 r = ld4(rstd + c4 * 4);
        const float4 gm = gamma ? ld4(gamma + c4 * 4) : f4(1.f);
        return State{ld4(mean + c4 * 4), r, mul4(gm, r), mul4(ld4(sg + c4 * 4), f4(inv_n)), mul4(ld4(sgx + c4 * 4), f4(inv_n)),
                     ld4(gprime + seg * C + c4 * 4), __ldg(reinterpret_cast<const int4*>(arg + seg * C + c4 * 4)),
                     seg * seg_rows};
    }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        const int rl = (int)(r - st.base);
        const float4 xh = mul4(sub4(ld4(x + i), st.mean), st.rstd);
        const float4 gi = make_float4(rl == st.arg.x ? st.g.x : 0.f, rl == st.arg.y ? st.g.y : 0.f,
                                      rl == st.arg.z ? st.g.z : 0.f, rl == st.arg.w ? st.g.w : 0.f);
        st4(dx + i, mul4(st.coef, sub4(sub4(gi, st.a), mul4(xh, st.b))));
    }
};

struct PoolLayout { Plan p; size_t off_max, off_amax, off_min, off_amin, total; };

inline PoolLayout pool_layout(int64_t R, int C, int64_t seg_rows) {
    PoolLayout l;
    l.p = make_plan(R, C / 4, PTX, seg_rows, 8 * kNumSMs);
    const size_t nby = (size_t)l.p.nseg * l.p.chunks;
    const size_t row = (size_t)C * sizeof(float);
    l.off_max = nby * 2 * row;
    l.off_amax = l.off_max + nby * row;
    l.off_min = l.off_amax + nby * row;
    l.off_amin = l.off_min + nby * row;
    l.total = l.off_amin + nby * row;
    return l;
}

}  // namespace

extern "C" size_t spgan_bn_pool_workspace(int64_t R, int C, int64_t seg_rows) {
    if (R <= 0 || C <= 0 || C % 4 != 0 || seg_rows <= 0 || R % seg_rows != 0) return 0;
    return pool_layout(R, C, seg_rows).total;
}

extern "C" int spgan_bn_pool_fwd(const float* x, int64_t R, int C, int64_t seg_rows, const float* gamma,
                                 const float* beta, float eps, float slope, float* mean, float* rstd, float* var,
                                 float* pooled, int32_t* arg, void* workspace, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && mean && rstd && pooled && arg && workspace && R >= 1 && C >= 4 && seg_rows >= 1);
    SPGAN_CHECK_ARG(R % seg_rows == 0 && slope > 0.f);
    if (C % 4 != 0 || !al16(x) || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return SPGAN_E_UNSUPPORTED;
    const PoolLayout l = pool_layout(R, C, seg_rows);
    const int64_t gy = l.p.nseg * l.p.chunks;
    if (gy > 65535 || seg_rows > 0x7fffffffLL) return SPGAN_E_UNSUPPORTED;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    float* psum = reinterpret_cast<float*>(ws);
    float* pmax = reinterpret_cast<float*>(ws + l.off_max);
    int* pamax = reinterpret_cast<int*>(ws + l.off_amax);
    float* pmin = reinterpret_cast<float*>(ws + l.off_min);
    int* pamin = reinterpret_cast<int*>(ws + l.off_amin);
    const int C4 = C / 4;
    dim3 grid((C4 + PTX - 1) / PTX, (unsigned)gy);
    bn_pool_partial_kernel<<<grid, PTX * PTY, 0, as_stream(s)>>>(x, C4, seg_rows, l.p.chunks, l.p.rows_per_chunk, psum,
                                                                 pmax, pamax, pmin, pamin);
    int rc = spgan_launch_status();
    if (rc != SPGAN_OK) return rc;
    bn_pool_final_kernel<<<(C + 31) / 32, 256, 0, as_stream(s)>>>(x, C, l.p.nseg, seg_rows, l.p.chunks, eps, psum, pmax,
                                                                  pamax, pmin, pamin, gamma, beta, slope, mean, rstd,
                                                                  var, pooled, arg);
    return spgan_launch_status();
}

extern "C" int spgan_bn_pool_bwd(const float* gpooled, const float* x, const int32_t* arg, int64_t R, int C,
                                 int64_t seg_rows, const float* mean, const float* rstd, const float* gamma,
                                 const float* beta, float slope, float* gprime /*[nseg,C] scratch*/, float* sg,
                                 float* sgx, float* dx, spgan_stream_t s) {
    SPGAN_CHECK_ARG(gpooled && x && arg && mean && rstd && gprime && sg && sgx && R >= 1 && C >= 4 && seg_rows >= 1);
    SPGAN_CHECK_ARG(R % seg_rows == 0);
    if (C % 4 != 0 || !al16(x) || (dx && !al16(dx)) || !al16(gprime) || !al16(arg)) return SPGAN_E_UNSUPPORTED;
    const int64_t nseg = R / seg_rows;
    bn_pool_bwd_sums_kernel<<<(C + 31) / 32, 256, 0, as_stream(s)>>>(gpooled, x, arg, C, nseg, seg_rows, mean, rstd, gamma,
                                                                     beta, slope, gprime, sg, sgx);
    int rc = spgan_launch_status();
    if (rc != SPGAN_OK || dx == nullptr) return rc;
    return run_map(R, C, seg_rows, as_stream(s),
                   BnPoolBwdApplyOp4{x, C, seg_rows, 1.f / (float)R, mean, rstd, gamma, sg, sgx, gprime, arg, dx});
}
