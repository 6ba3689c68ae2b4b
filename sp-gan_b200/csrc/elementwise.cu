// Elementwise, broadcast and layout kernels (HBM-bound; float4 where alignment allows).
#include "common.cuh"

namespace {

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Generic flat map: f(i) for scalars, f4(i4) for aligned float4 groups.
template <typename F>
__global__ void map1_kernel(int64_t n, F f) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i);
}

template <typename F>
int launch_map(int64_t n, cudaStream_t st, F f) {
    if (n <= 0) return SPGAN_OK;
    map1_kernel<<<ew_grid(n, 256, 16), 256, 0, st>>>(n, f);
    return spgan_launch_status();
}

// unary / binary ops with a float4 fast path
template <typename Op>
__global__ void ew_vec_kernel(int64_t n4, Op op) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        op.vec(i);
}
template <typename Op>
__global__ void ew_tail_kernel(int64_t beg, int64_t n, Op op) {
    for (int64_t i = beg + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        op.scalar(i);
}
template <typename Op>
int launch_ew(int64_t n, bool can_vec, cudaStream_t st, Op op) {
    if (n <= 0) return SPGAN_OK;
    int64_t n4 = can_vec ? n / 4 : 0;
    if (n4 > 0) ew_vec_kernel<<<ew_grid(n4, 256, 16), 256, 0, st>>>(n4, op);
    if (n4 * 4 < n) ew_tail_kernel<<<ew_grid(n - n4 * 4, 256, 16), 256, 0, st>>>(n4 * 4, n, op);
    return spgan_launch_status();
}

#define F4(p) reinterpret_cast<float4*>(p)
#define CF4(p) reinterpret_cast<const float4*>(p)

struct FillOp {
    float* x; float v;
    __device__ void vec(int64_t i) const { F4(x)[i] = make_float4(v, v, v, v); }
    __device__ void scalar(int64_t i) const { x[i] = v; }
};
struct CopyOp {
    const float* x; float* y;
    __device__ void vec(int64_t i) const { F4(y)[i] = __ldg(CF4(x) + i); }
    __device__ void scalar(int64_t i) const { y[i] = x[i]; }
};
struct AxpbyOp {
    float a; const float* x; float b; const float* y; float* o;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i), w = __ldg(CF4(y) + i);
        F4(o)[i] = make_float4(a * u.x + b * w.x, a * u.y + b * w.y, a * u.z + b * w.z, a * u.w + b * w.w);
    }
    __device__ void scalar(int64_t i) const { o[i] = a * x[i] + b * y[i]; }
};
struct ScaleOp {   // y == nullptr variant of axpby
    float a; const float* x; float* o;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i);
        F4(o)[i] = make_float4(a * u.x, a * u.y, a * u.z, a * u.w);
    }
    __device__ void scalar(int64_t i) const { o[i] = a * x[i]; }
};
struct MulOp {
    const float* x; const float* y; float* o;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i), w = __ldg(CF4(y) + i);
        F4(o)[i] = make_float4(u.x * w.x, u.y * w.y, u.z * w.z, u.w * w.w);
    }
    __device__ void scalar(int64_t i) const { o[i] = x[i] * y[i]; }
};
struct LreluOp {
    const float* x; float s; float* y;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i);
        F4(y)[i] = make_float4(lrelu_f(u.x, s), lrelu_f(u.y, s), lrelu_f(u.z, s), lrelu_f(u.w, s));
    }
    __device__ void scalar(int64_t i) const { y[i] = lrelu_f(x[i], s); }
};
struct LreluBwdOp {
    const float* g; const float* x; float s; float* d;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i), w = __ldg(CF4(g) + i);
        F4(d)[i] = make_float4(u.x > 0.f ? w.x : w.x * s, u.y > 0.f ? w.y : w.y * s, u.z > 0.f ? w.z : w.z * s,
                               u.w > 0.f ? w.w : w.w * s);
    }
    __device__ void scalar(int64_t i) const { d[i] = x[i] > 0.f ? g[i] : g[i] * s; }
};
struct TanhOp {
    const float* x; float* y;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(x) + i);
        F4(y)[i] = make_float4(tanhf(u.x), tanhf(u.y), tanhf(u.z), tanhf(u.w));
    }
    __device__ void scalar(int64_t i) const { y[i] = tanhf(x[i]); }
};
struct TanhBwdOp {
    const float* g; const float* y; float* d;
    __device__ void vec(int64_t i) const {
        float4 u = __ldg(CF4(y) + i), w = __ldg(CF4(g) + i);
        F4(d)[i] = make_float4(w.x * (1.f - u.x * u.x), w.y * (1.f - u.y * u.y), w.z * (1.f - u.z * u.z),
                               w.w * (1.f - u.w * u.w));
    }
    __device__ void scalar(int64_t i) const { d[i] = g[i] * (1.f - y[i] * y[i]); }
};

// out[r,c] = x[r,c] (op) v[r / seg_rows, c]
template <int MODE>   // 0 add, 1 mul, 2 broadcast only
__global__ void segvec_kernel(const float* __restrict__ x, const float* __restrict__ v, int64_t R, int C,
                              int64_t seg_rows, float* __restrict__ out) {
    const int64_t total = R * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const float w = __ldg(v + (r / seg_rows) * C + c);
        out[i] = MODE == 0 ? x[i] + w : (MODE == 1 ? x[i] * w : w);
    }
}
template <int MODE>
__global__ void segvec4_kernel(const float* __restrict__ x, const float* __restrict__ v, int64_t R, int C4,
                               int64_t seg_rows, float* __restrict__ out) {
    const int64_t total = R * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C4;
        const int c = (int)(i - r * C4);
        const float4 w = __ldg(CF4(v) + (r / seg_rows) * C4 + c);
        float4 o;
        if (MODE == 2) o = w;
        else {
            const float4 u = __ldg(CF4(x) + i);
            o = MODE == 0 ? make_float4(u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w)
                          : make_float4(u.x * w.x, u.y * w.y, u.z * w.z, u.w * w.w);
        }
        F4(out)[i] = o;
    }
}
template <int MODE>
int launch_segvec(const float* x, const float* v, int64_t R, int C, int64_t seg_rows, float* out, cudaStream_t st) {
    if (R <= 0) return SPGAN_OK;
    if (C % 4 == 0 && al16(v) && al16(out) && (MODE == 2 || al16(x))) {
        const int64_t total = R * (C / 4);
        segvec4_kernel<MODE><<<ew_grid(total, 256, 16), 256, 0, st>>>(x, v, R, C / 4, seg_rows, out);
    } else {
        const int64_t total = R * C;
        segvec_kernel<MODE><<<ew_grid(total, 256, 16), 256, 0, st>>>(x, v, R, C, seg_rows, out);
    }
    return spgan_launch_status();
}

// [B,C,N] (strided) -> rows [B*N, C] through a 32x32 shared tile
__global__ void bcn_to_rows_kernel(const float* __restrict__ src, int64_t sb, int64_t sc, int64_t sn, int C, int N,
                                   float* __restrict__ rows) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const float* sbp = src + (int64_t)b * sb;
    if (sn <= sc) {   // points contiguous-ish: read along n
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r, n = n0 + tx;
            tile[r][tx] = (c < C && n < N) ? __ldg(sbp + (int64_t)c * sc + (int64_t)n * sn) : 0.f;
        }
    } else {          // channels contiguous-ish (e.g. a transposed view of [B,N,3]): read along c
        for (int r = ty; r < 32; r += 8) {
            const int n = n0 + r, c = c0 + tx;
            tile[tx][r] = (c < C && n < N) ? __ldg(sbp + (int64_t)c * sc + (int64_t)n * sn) : 0.f;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, c = c0 + tx;
        if (c < C && n < N) rows[((int64_t)b * N + n) * C + c] = tile[tx][r];
    }
}

__global__ void rows_to_bcn_kernel(const float* __restrict__ rows, int C, int N, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, c = c0 + tx;
        tile[r][tx] = (c < C && n < N) ? __ldg(rows + ((int64_t)b * N + n) * C + c) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, n = n0 + tx;
        if (c < C && n < N) dst[((int64_t)b * C + c) * N + n] = tile[tx][r];
    }
}

__global__ void concat_cols_kernel(const float* __restrict__ a, int64_t lda, int Ca, const float* __restrict__ b,
                                   int64_t ldb, int64_t b_seg_stride, int seg_rows, int Cb, int64_t R,
                                   float* __restrict__ out) {
    const int Cc = Ca + Cb;
    const int64_t total = R * Cc;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Cc;
        const int c = (int)(i - r * Cc);
        float v;
        if (c < Ca) v = __ldg(a + r * lda + c);
        else {
            const int64_t seg = r / seg_rows, rr = r - seg * seg_rows;
            v = __ldg(b + seg * b_seg_stride + rr * ldb + (c - Ca));
        }
        out[i] = v;
    }
}

__global__ void split_cols_kernel(const float* __restrict__ g, int64_t R, int Ca, int Cb, float* __restrict__ ga,
                                  float* __restrict__ gb) {
    const int Cc = Ca + Cb;
    const int64_t total = R * Cc;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Cc;
        const int c = (int)(i - r * Cc);
        const float v = g[i];
        if (c < Ca) { if (ga) ga[r * Ca + c] = v; }
        else if (gb) gb[r * Cb + (c - Ca)] = v;
    }
}

__global__ void permute_ock_okc_kernel(const float* __restrict__ src, int O, int Cc, int k, float* __restrict__ dst,
                                       bool inverse) {
    const int64_t total = (int64_t)O * Cc * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates the [O, Cc, k] layout
        const int r = (int)(i % k);
        const int c = (int)((i / k) % Cc);
        const int64_t o = i / ((int64_t)k * Cc);
        const int64_t j = (o * k + r) * Cc + c;   // [O, k, Cc]
        if (!inverse) dst[j] = src[i]; else dst[i] = src[j];
    }
}

__global__ void gp_interp_kernel(const float* __restrict__ real, int64_t rsb, int64_t rsc, int64_t rsn,
                                 const float* __restrict__ fake, int64_t fsb, int64_t fsc, int64_t fsn,
                                 const float* __restrict__ alpha, int B, int C, int N, float* __restrict__ mix) {
    const int64_t total = (int64_t)B * C * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int c = (int)((i / N) % C);
        const int b = (int)(i / ((int64_t)N * C));
        const float r = __ldg(real + b * rsb + c * rsc + n * rsn);
        const float f = __ldg(fake + b * fsb + c * fsc + n * fsn);
        mix[i] = r + __ldg(alpha + b) * (f - r);
    }
}

__global__ void rsqrt_eps_kernel(const float* __restrict__ v, float eps, int64_t n, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = 1.f / sqrtf(v[i] + eps);
}

// one warp per row
__global__ void row_l2_normalize_kernel(const float* __restrict__ x, int64_t R, int C, float eps,
                                        float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = __ldg(x + r * C + c); s = fmaf(v, v, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = 1.f / (sqrtf(s) + eps);
        for (int c = lane; c < C; c += 32) out[r * C + c] = __ldg(x + r * C + c) * inv;
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float gscale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

// Device-resident step counter variant (CUDA-graph replay: nothing that changes between steps may be a kernel
// argument).  state = {int32 step, float 1 - beta1^step, float sqrt(1 - beta2^step), pad}.
__global__ void adam_prep_kernel(int32_t* state, float b1, float b2) {
    const int step = state[0] + 1;
    state[0] = step;
    float* f = reinterpret_cast<float*>(state);
    f[1] = 1.f - powf(b1, (float)step);
    f[2] = sqrtf(1.f - powf(b2, (float)step));
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                const int32_t* __restrict__ state, float gscale) {
    const float bc1 = reinterpret_cast<const float*>(state)[1], bc2_sqrt = reinterpret_cast<const float*>(state)[2];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
    }
}

}  // namespace

extern "C" int spgan_fill(float* x, int64_t n, float v, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && (x || n == 0));
    return launch_ew(n, al16(x), as_stream(s), FillOp{x, v});
}
extern "C" int spgan_copy(const float* x, float* y, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((x && y) || n == 0));
    return launch_ew(n, al16(x) && al16(y), as_stream(s), CopyOp{x, y});
}
extern "C" int spgan_axpby(float a, const float* x, float b, const float* y, float* out, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((x && out) || n == 0));
    if (y == nullptr) return launch_ew(n, al16(x) && al16(out), as_stream(s), ScaleOp{a, x, out});
    return launch_ew(n, al16(x) && al16(y) && al16(out), as_stream(s), AxpbyOp{a, x, b, y, out});
}
extern "C" int spgan_mul(const float* x, const float* y, float* out, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((x && y && out) || n == 0));
    return launch_ew(n, al16(x) && al16(y) && al16(out), as_stream(s), MulOp{x, y, out});
}
extern "C" int spgan_lrelu(const float* x, float slope, float* y, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((x && y) || n == 0));
    return launch_ew(n, al16(x) && al16(y), as_stream(s), LreluOp{x, slope, y});
}
extern "C" int spgan_lrelu_bwd(const float* g, const float* x, float slope, float* dx, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((g && x && dx) || n == 0));
    return launch_ew(n, al16(g) && al16(x) && al16(dx), as_stream(s), LreluBwdOp{g, x, slope, dx});
}
extern "C" int spgan_tanh(const float* x, float* y, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((x && y) || n == 0));
    return launch_ew(n, al16(x) && al16(y), as_stream(s), TanhOp{x, y});
}
extern "C" int spgan_tanh_bwd(const float* g, const float* y, float* dx, int64_t n, spgan_stream_t s) {
    SPGAN_CHECK_ARG(n >= 0 && ((g && y && dx) || n == 0));
    return launch_ew(n, al16(g) && al16(y) && al16(dx), as_stream(s), TanhBwdOp{g, y, dx});
}
extern "C" int spgan_add_segvec(const float* x, const float* v, int64_t R, int C, int64_t seg_rows, float* out,
                                spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && v && out && R >= 0 && C >= 1 && seg_rows >= 1);
    return launch_segvec<0>(x, v, R, C, seg_rows, out, as_stream(s));
}
extern "C" int spgan_mul_segvec(const float* x, const float* v, int64_t R, int C, int64_t seg_rows, float* out,
                                spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && v && out && R >= 0 && C >= 1 && seg_rows >= 1);
    return launch_segvec<1>(x, v, R, C, seg_rows, out, as_stream(s));
}
extern "C" int spgan_bcast_segvec(const float* v, int64_t R, int C, int64_t seg_rows, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(v && out && R >= 0 && C >= 1 && seg_rows >= 1);
    return launch_segvec<2>(nullptr, v, R, C, seg_rows, out, as_stream(s));
}

extern "C" int spgan_bcn_to_rows(const float* src, int64_t sb, int64_t sc, int64_t sn, int B, int C, int N,
                                 float* rows, spgan_stream_t s) {
    SPGAN_CHECK_ARG(src && rows && B >= 0 && C >= 1 && N >= 1);
    if (B == 0) return SPGAN_OK;
    if (B > 65535 || (C + 31) / 32 > 65535) return SPGAN_E_UNSUPPORTED;
    dim3 grid((N + 31) / 32, (C + 31) / 32, B), block(32, 8);
    bcn_to_rows_kernel<<<grid, block, 0, as_stream(s)>>>(src, sb, sc, sn, C, N, rows);
    return spgan_launch_status();
}
extern "C" int spgan_rows_to_bcn(const float* rows, int B, int C, int N, float* dst, spgan_stream_t s) {
    SPGAN_CHECK_ARG(rows && dst && B >= 0 && C >= 1 && N >= 1);
    if (B == 0) return SPGAN_OK;
    if (B > 65535 || (C + 31) / 32 > 65535) return SPGAN_E_UNSUPPORTED;
    dim3 grid((N + 31) / 32, (C + 31) / 32, B), block(32, 8);
    rows_to_bcn_kernel<<<grid, block, 0, as_stream(s)>>>(rows, C, N, dst);
    return spgan_launch_status();
}
extern "C" int spgan_concat_cols(const float* a, int64_t lda, int Ca, const float* b, int64_t ldb,
                                 int64_t b_seg_stride, int seg_rows, int Cb, int64_t R, float* out,
                                 spgan_stream_t s) {
    SPGAN_CHECK_ARG(a && b && out && Ca >= 1 && Cb >= 1 && R >= 0 && seg_rows >= 1);
    if (R == 0) return SPGAN_OK;
    concat_cols_kernel<<<ew_grid(R * (Ca + Cb), 256, 16), 256, 0, as_stream(s)>>>(a, lda, Ca, b, ldb, b_seg_stride,
                                                                                  seg_rows, Cb, R, out);
    return spgan_launch_status();
}
extern "C" int spgan_split_cols_add(const float* g, int64_t R, int Ca, int Cb, float* ga, float* gb,
                                    spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && Ca >= 1 && Cb >= 1 && R >= 0);
    if (R == 0) return SPGAN_OK;
    split_cols_kernel<<<ew_grid(R * (Ca + Cb), 256, 16), 256, 0, as_stream(s)>>>(g, R, Ca, Cb, ga, gb);
    return spgan_launch_status();
}
extern "C" int spgan_permute_ock_to_okc(const float* src, int O, int Cc, int k, float* dst, spgan_stream_t s) {
    SPGAN_CHECK_ARG(src && dst && O >= 1 && Cc >= 1 && k >= 1);
    permute_ock_okc_kernel<<<ew_grid((int64_t)O * Cc * k, 256), 256, 0, as_stream(s)>>>(src, O, Cc, k, dst, false);
    return spgan_launch_status();
}
extern "C" int spgan_permute_okc_to_ock(const float* src, int O, int Cc, int k, float* dst, spgan_stream_t s) {
    SPGAN_CHECK_ARG(src && dst && O >= 1 && Cc >= 1 && k >= 1);
    permute_ock_okc_kernel<<<ew_grid((int64_t)O * Cc * k, 256), 256, 0, as_stream(s)>>>(src, O, Cc, k, dst, true);
    return spgan_launch_status();
}
extern "C" int spgan_gp_interp(const float* real, int64_t rsb, int64_t rsc, int64_t rsn, const float* fake,
                               int64_t fsb, int64_t fsc, int64_t fsn, const float* alpha, int B, int C, int N,
                               float* mix, spgan_stream_t s) {
    SPGAN_CHECK_ARG(real && fake && alpha && mix && B >= 0 && C >= 1 && N >= 1);
    if (B == 0) return SPGAN_OK;
    gp_interp_kernel<<<ew_grid((int64_t)B * C * N, 256), 256, 0, as_stream(s)>>>(real, rsb, rsc, rsn, fake, fsb, fsc,
                                                                               fsn, alpha, B, C, N, mix);
    return spgan_launch_status();
}
extern "C" int spgan_rsqrt_eps(const float* v, float eps, int64_t n, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(v && out && n >= 0);
    if (n == 0) return SPGAN_OK;
    rsqrt_eps_kernel<<<ew_grid(n, 256), 256, 0, as_stream(s)>>>(v, eps, n, out);
    return spgan_launch_status();
}
extern "C" int spgan_row_l2_normalize(const float* x, int64_t R, int C, float eps, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && out && R >= 0 && C >= 1);
    if (R == 0) return SPGAN_OK;
    row_l2_normalize_kernel<<<ew_grid(R * 32, 256), 256, 0, as_stream(s)>>>(x, R, C, eps, out);
    return spgan_launch_status();
}
extern "C" int spgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                               float beta2, float eps, int step, float grad_scale, spgan_stream_t s) {
    SPGAN_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1);
    if (n == 0) return SPGAN_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    adam_kernel<<<ew_grid(n, 256), 256, 0, as_stream(s)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
    return spgan_launch_status();
}

extern "C" int spgan_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                   float beta2, float eps, int32_t* state, float grad_scale, spgan_stream_t s) {
    SPGAN_CHECK_ARG(p && g && m && v && state && n >= 0);
    adam_prep_kernel<<<1, 1, 0, as_stream(s)>>>(state, beta1, beta2);
    if (n == 0) return spgan_launch_status();
    adam_dev_kernel<<<ew_grid(n, 256), 256, 0, as_stream(s)>>>(p, g, m, v, n, lr, beta1, beta2, eps, state, grad_scale);
    return spgan_launch_status();
}

extern "C" int spgan_abi_version(void) { return SPGAN_ABI_VERSION; }
extern "C" const char* spgan_error_string(int code) {
    switch (code) {
        case SPGAN_OK: return "ok";
        case SPGAN_E_BADARG: return "bad argument (null pointer, non-positive size or inconsistent shape)";
        case SPGAN_E_UNSUPPORTED: return "request outside the implemented envelope";
        case SPGAN_E_ALIGN: return "pointer or leading dimension violates the required alignment";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown spgan error";
    }
}
