// Column reductions, batch/instance normalisation (forward, backward, double backward),
// poolings over points / neighbours, softmax over neighbours, edge aggregation, penalty.
#include "common.cuh"
#include <stdlib.h>
#include "norm_fast.cuh"
#include <float.h>

namespace {

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------
// generic segmented column reduction: partial sums per (segment, row-chunk), then a
// double-precision finalize.  Block = 32 columns x 8 row lanes.
// ---------------------------------------------------------------------------------------
struct ChunkPlan { int64_t nseg; int chunks; int64_t rows_per_chunk; };

inline ChunkPlan plan_chunks(int64_t R, int C, int64_t seg_rows) {
    ChunkPlan p;
    p.nseg = R / seg_rows;
    const int64_t col_blocks = (C + 31) / 32;
    int64_t want = (8LL * kNumSMs + col_blocks * p.nseg - 1) / (col_blocks * p.nseg);
    int64_t maxc = (seg_rows + 63) / 64;
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    p.rows_per_chunk = ((seg_rows + want - 1) / want + 7) / 8 * 8;
    p.chunks = (int)((seg_rows + p.rows_per_chunk - 1) / p.rows_per_chunk);
    return p;
}

template <int NV, typename Op>
__global__ void __launch_bounds__(256)
colreduce_partial_kernel(Op op, int C, int64_t seg_rows, int chunks, int64_t rows_per_chunk,
                         float* __restrict__ partial) {
    __shared__ float red[8][NV][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const int64_t seg = blockIdx.y / chunks;
    const int chunk = blockIdx.y % chunks;
    const int64_t rbeg = seg * seg_rows + (int64_t)chunk * rows_per_chunk;
    int64_t rend = rbeg + rows_per_chunk;
    const int64_t seg_end = (seg + 1) * seg_rows;
    if (rend > seg_end) rend = seg_end;
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int64_t r = rbeg + ty; r < rend; r += 8) op(r, c, seg, acc);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) red[ty][v][tx] = acc[v];
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) s += red[t][v][tx];
            partial[((int64_t)blockIdx.y * NV + v) * C + c] = s;
        }
    }
}

constexpr int FIN_TX = 8;                 // columns per finalize block
constexpr int FIN_TY = 64;                // chunk lanes of the finalize block
template <int NV, typename Fin>
__global__ void __launch_bounds__(FIN_TX * FIN_TY)
colreduce_final_kernel(const float* __restrict__ partial, int C, int64_t nseg, int chunks, Fin fin) {
    // block = 8 columns x 64 chunk lanes of one segment.  A narrow tensor has up to ~1200 chunk rows and few
    // columns, so the pass is latency-bound: narrow column groups give C / 8 blocks (a 32-column block left a C = 64
    // BatchNorm with 2 blocks walking 73 dependent batches), every thread keeps 8 independent loads in flight, and
    // each partial row is read as one 32-byte sector per group.  Double-precision combine in a fixed order =>
    // deterministic.
    __shared__ double red[FIN_TY][NV][FIN_TX + 1];
    const int tx = threadIdx.x % FIN_TX, ty = threadIdx.x / FIN_TX;
    const int c = blockIdx.x * FIN_TX + tx;
    for (int64_t seg = blockIdx.y; seg < nseg; seg += gridDim.y) {
        double s[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] = 0.0;
        if (c < C) {
            const float* base = partial + (seg * chunks) * NV * (int64_t)C + c;
            int ch = ty;
            for (; ch + 7 * FIN_TY < chunks; ch += 8 * FIN_TY) {
                float t[8][NV];
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int v = 0; v < NV; ++v) t[q][v] = __ldg(base + ((int64_t)(ch + q * FIN_TY) * NV + v) * C);
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int v = 0; v < NV; ++v) s[v] += (double)t[q][v];
            }
            {   // tail: up to 7 more rows per lane, again all loads first
                float t[7][NV];
#pragma unroll
                for (int q = 0; q < 7; ++q)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        t[q][v] = (ch + q * FIN_TY < chunks) ? __ldg(base + ((int64_t)(ch + q * FIN_TY) * NV + v) * C) : 0.f;
#pragma unroll
                for (int q = 0; q < 7; ++q)
#pragma unroll
                    for (int v = 0; v < NV; ++v) s[v] += (double)t[q][v];
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) red[ty][v][tx] = s[v];
        __syncthreads();
        // two fixed-order levels: lanes ty < 8 fold rows ty*8 .. ty*8+7, lane 0 folds the 8 results
        if (ty < 8) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double t = red[ty * 8][v][tx];
#pragma unroll
                for (int y = 1; y < 8; ++y) t += red[ty * 8 + y][v][tx];
                s[v] = t;
            }
        }
        __syncthreads();
        if (ty < 8) {
#pragma unroll
            for (int v = 0; v < NV; ++v) red[ty][v][tx] = s[v];
        }
        __syncthreads();
        if (ty == 0 && c < C) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                double t = red[0][v][tx];
#pragma unroll
                for (int y = 1; y < 8; ++y) t += red[y][v][tx];
                s[v] = t;
            }
            fin(seg, c, s);
        }
        __syncthreads();
    }
}

// few chunks per segment (many short segments): one thread per (segment, column), coalesced over columns
template <int NV, typename Fin>
__global__ void colreduce_final_small_kernel(const float* __restrict__ partial, int C, int64_t nseg, int chunks, Fin fin) {
    const int64_t total = nseg * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t seg = i / C;
        const int c = (int)(i - seg * C);
        double s[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] = 0.0;
        for (int ch = 0; ch < chunks; ++ch)
#pragma unroll
            for (int v = 0; v < NV; ++v) s[v] += (double)__ldg(partial + (((seg * chunks) + ch) * NV + v) * C + c);
        fin(seg, c, s);
    }
}

// (measured: 24 * 148 blocks is much slower -- chunks of a few dozen rows make the per-block reduction and the
// finalize pass dominate; 8 * 148 stays)
inline int reduce_target_blocks() {       // SPGAN_REDUCE_WAVES = CTAs per SM the partial pass aims for (tuning knob)
    static const int v = [] {
        const char* e = getenv("SPGAN_REDUCE_WAVES");
        const int w = e ? atoi(e) : 0;
        return (w >= 1 && w <= 32 ? w : 8) * kNumSMs;
    }();
    return v;
}
#define kReduceTargetBlocks reduce_target_blocks()

template <int NV, typename Op, typename Op4, typename Fin>
int run_colreduce(int64_t R, int C, int64_t seg_rows, void* workspace, cudaStream_t st, Op op, Op4 op4, bool can_vec,
                  Fin fin) {
    if (R <= 0) return SPGAN_OK;
    if (seg_rows < 1 || R % seg_rows != 0 || workspace == nullptr) return SPGAN_E_BADARG;
    float* partial = reinterpret_cast<float*>(workspace);
    int64_t nseg;
    int chunks;
    if (can_vec && C % 4 == 0) {
        const int C4 = C / 4, TX = fastnorm::pick_tx(C4);
        const fastnorm::Plan p = fastnorm::make_plan(R, C4, TX, seg_rows, kReduceTargetBlocks);
        const int64_t gy = p.nseg * p.chunks;
        if (gy > 65535) return SPGAN_E_UNSUPPORTED;
        dim3 grid((C4 + TX - 1) / TX, (unsigned)gy);
        if (TX == 32) fastnorm::colreduce4_kernel<NV, 32><<<grid, 256, 0, st>>>(op4, C4, seg_rows, p.chunks, p.rows_per_chunk, partial);
        else if (TX == 16) fastnorm::colreduce4_kernel<NV, 16><<<grid, 256, 0, st>>>(op4, C4, seg_rows, p.chunks, p.rows_per_chunk, partial);
        else fastnorm::colreduce4_kernel<NV, 8><<<grid, 256, 0, st>>>(op4, C4, seg_rows, p.chunks, p.rows_per_chunk, partial);
        nseg = p.nseg; chunks = p.chunks;
    } else {
        const ChunkPlan p = plan_chunks(R, C, seg_rows);
        const int64_t gy = p.nseg * p.chunks;
        if (gy > 65535) return SPGAN_E_UNSUPPORTED;
        dim3 grid((C + 31) / 32, (unsigned)gy), block(32, 8);
        colreduce_partial_kernel<NV><<<grid, block, 0, st>>>(op, C, seg_rows, p.chunks, p.rows_per_chunk, partial);
        nseg = p.nseg; chunks = p.chunks;
    }
    if (chunks <= 8)
        colreduce_final_small_kernel<NV><<<ew_grid(nseg * C, 256), 256, 0, st>>>(partial, C, nseg, chunks, fin);
    else
        colreduce_final_kernel<NV><<<dim3((unsigned)((C + FIN_TX - 1) / FIN_TX), (unsigned)(nseg < 4096 ? nseg : 4096)), FIN_TX * FIN_TY, 0, st>>>(partial, C, nseg, chunks, fin);
    return spgan_launch_status();
}

struct SumOp {
    const float* x; int C;
    __device__ void operator()(int64_t r, int c, int64_t, float* acc) const { acc[0] += __ldg(x + r * C + c); }
};
struct DotOp {
    const float* x; const float* y; int C;
    __device__ void operator()(int64_t r, int c, int64_t, float* acc) const {
        acc[0] += __ldg(x + r * C + c) * __ldg(y + r * C + c);
    }
};
struct Store1Fin {
    float* out; int C;
    __device__ void operator()(int64_t seg, int c, const double* s) const { out[seg * C + c] = (float)s[0]; }
};
// shifted moments: d = x - x[first row of segment]
struct StatsOp {
    const float* x; int C; int64_t seg_rows;
    __device__ void operator()(int64_t r, int c, int64_t seg, float* acc) const {
        const float d = __ldg(x + r * C + c) - __ldg(x + seg * seg_rows * C + c);
        acc[0] += d; acc[1] = fmaf(d, d, acc[1]);
    }
};
struct StatsFin {
    const float* x; int C; int64_t seg_rows; float eps; float* mean; float* rstd; float* var;
    // optional (one segment only): nn.BatchNorm's running-statistics update folded into the finalize
    float momentum; float unbias; float* rm; float* rv; int64_t* count;
    __device__ void operator()(int64_t seg, int c, const double* s) const {
        const double n = (double)seg_rows;
        const double m1 = s[0] / n;
        double v = s[1] / n - m1 * m1;
        if (v < 0.0) v = 0.0;
        const int64_t o = seg * C + c;
        const float m = (float)((double)x[seg * seg_rows * C + c] + m1);
        mean[o] = m;
        rstd[o] = (float)(1.0 / sqrt(v + (double)eps));
        if (var) var[o] = (float)v;
        if (rm) {
            rm[c] = (1.f - momentum) * rm[c] + momentum * m;
            rv[c] = (1.f - momentum) * rv[c] + momentum * ((float)v * unbias);
            if (c == 0 && count) *count += 1;
        }
    }
};
// The LeakyReLU mask of the fused BN+activation is recomputed from x (same fma as the forward) instead of
// re-reading the activated output: one fewer full-tensor read per pass.
struct NormBwdOp {
    const float* g; const float* x; float slope; int C; const float* mean; const float* rstd;
    const float* gamma; const float* beta;
    __device__ void operator()(int64_t r, int c, int64_t seg, float* acc) const {
        const int64_t i = r * C + c;
        float gi = __ldg(g + i);
        const float xh = (__ldg(x + i) - __ldg(mean + seg * C + c)) * __ldg(rstd + seg * C + c);
        if (slope != 1.f) {
            const float pre = fmaf(xh, gamma ? __ldg(gamma + c) : 1.f, beta ? __ldg(beta + c) : 0.f);
            if (!(pre > 0.f)) gi *= slope;
        }
        acc[0] += gi; acc[1] = fmaf(gi, xh, acc[1]);
    }
};
struct Store2Fin {
    float* a; float* b; int C;
    // optional (one segment only): accumulate into the parameter gradients (dbeta += sum g', dgamma += sum g' xhat)
    float* acc_a; float* acc_b;
    __device__ void operator()(int64_t seg, int c, const double* s) const {
        a[seg * C + c] = (float)s[0]; b[seg * C + c] = (float)s[1];
        if (acc_a) acc_a[c] += (float)s[0];
        if (acc_b) acc_b[c] += (float)s[1];
    }
};
struct DblBwdOp {
    const float* g; const float* u; const float* x; int C; const float* mean;
    __device__ void operator()(int64_t r, int c, int64_t, float* acc) const {
        const int64_t i = r * C + c;
        const float gi = __ldg(g + i), ui = __ldg(u + i), xm = __ldg(x + i) - __ldg(mean + c);
        acc[0] += gi; acc[1] += ui; acc[2] = fmaf(gi, xm, acc[2]); acc[3] = fmaf(ui, xm, acc[3]);
        acc[4] = fmaf(gi, ui, acc[4]);
    }
};
struct Store5Fin {
    float* out; int C;
    __device__ void operator()(int64_t, int c, const double* s) const {
#pragma unroll
        for (int v = 0; v < 5; ++v) out[v * C + c] = (float)s[v];
    }
};

// ---- float4 variants (see norm_fast.cuh): per-column parameters hoisted into State
using fastnorm::ld4; using fastnorm::st4; using fastnorm::f4; using fastnorm::add4; using fastnorm::sub4;
using fastnorm::mul4; using fastnorm::fma4; using fastnorm::mask4; using fastnorm::lrelu4;
struct NoState {};
struct SumOp4 {
    static constexpr int kStreams = 1;      // input tensors read per row (sets the unroll depth)
    const float* x; int C;
    using State = NoState;
    __device__ State init(int, int64_t) const { return State{}; }
    __device__ void accum(const State&, int64_t r, int c4, float4* acc) const { acc[0] = add4(acc[0], ld4(x + r * C + c4 * 4)); }
};
struct DotOp4 {
    static constexpr int kStreams = 2;      // input tensors read per row (sets the unroll depth)
    const float* x; const float* y; int C;
    using State = NoState;
    __device__ State init(int, int64_t) const { return State{}; }
    __device__ void accum(const State&, int64_t r, int c4, float4* acc) const {
        acc[0] = fma4(ld4(x + r * C + c4 * 4), ld4(y + r * C + c4 * 4), acc[0]);
    }
};
struct StatsOp4 {
    static constexpr int kStreams = 1;      // input tensors read per row (sets the unroll depth)
    const float* x; int C; int64_t seg_rows;
    struct State { float4 shift; };
    __device__ State init(int c4, int64_t seg) const { return State{ld4(x + seg * seg_rows * C + c4 * 4)}; }
    __device__ void accum(const State& st, int64_t r, int c4, float4* acc) const {
        const float4 d = sub4(ld4(x + r * C + c4 * 4), st.shift);
        acc[0] = add4(acc[0], d); acc[1] = fma4(d, d, acc[1]);
    }
};
struct NormBwdOp4 {
    static constexpr int kStreams = 2;      // input tensors read per row (sets the unroll depth)
    const float* g; const float* x; float slope; int C; const float* mean; const float* rstd;
    const float* gamma; const float* beta;
    struct State { float4 mean, rstd, gamma, beta; };
    __device__ State init(int c4, int64_t seg) const {
        return State{ld4(mean + seg * C + c4 * 4), ld4(rstd + seg * C + c4 * 4),
                     gamma ? ld4(gamma + c4 * 4) : f4(1.f), beta ? ld4(beta + c4 * 4) : f4(0.f)};
    }
    __device__ void accum(const State& st, int64_t r, int c4, float4* acc) const {
        const int64_t i = r * C + c4 * 4;
        float4 gi = ld4(g + i);
        const float4 xh = mul4(sub4(ld4(x + i), st.mean), st.rstd);
        if (slope != 1.f) gi = mask4(gi, fma4(xh, st.gamma, st.beta), slope);
        acc[0] = add4(acc[0], gi); acc[1] = fma4(gi, xh, acc[1]);
    }
};
struct DblBwdOp4 {
    static constexpr int kStreams = 3;      // input tensors read per row (sets the unroll depth)
    const float* g; const float* u; const float* x; int C; const float* mean;
    struct State { float4 mean; };
    __device__ State init(int c4, int64_t) const { return State{ld4(mean + c4 * 4)}; }
    __device__ void accum(const State& st, int64_t r, int c4, float4* acc) const {
        const int64_t i = r * C + c4 * 4;
        const float4 gi = ld4(g + i), ui = ld4(u + i), xm = sub4(ld4(x + i), st.mean);
        acc[0] = add4(acc[0], gi); acc[1] = add4(acc[1], ui); acc[2] = fma4(gi, xm, acc[2]);
        acc[3] = fma4(ui, xm, acc[3]); acc[4] = fma4(gi, ui, acc[4]);
    }
};
// maps
struct NormApplyOp4 {
    const float* x; int C; const float* mean; const float* rstd; const float* gamma; const float* beta; float slope; float* y;
    struct State { float4 mean, rstd, gamma, beta; };
    __device__ State init(int c4, int64_t seg) const {
        return State{ld4(mean + seg * C + c4 * 4), ld4(rstd + seg * C + c4 * 4),
                     gamma ? ld4(gamma + c4 * 4) : f4(1.f), beta ? ld4(beta + c4 * 4) : f4(0.f)};
    }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        float4 v = fma4(mul4(sub4(ld4(x + i), st.mean), st.rstd), st.gamma, st.beta);
        if (slope != 1.f) v = lrelu4(v, slope);
        st4(y + i, v);
    }
};
struct NormBwdApplyOp4 {
    const float* g; const float* x; float slope; int C; float inv_n; const float* mean;
    const float* rstd; const float* gamma; const float* beta; const float* sg; const float* sgx; float* dx;
    struct State { float4 mean, rstd, gamma, beta, coef, a, b; };     // dx = coef * (g' - a - xh * b)
    __device__ State init(int c4, int64_t seg) const {
        const float4 m = ld4(mean + seg * C + c4 * 4), r = ld4(rstd + seg * C + c4 * 4);
        const float4 gm = gamma ? ld4(gamma + c4 * 4) : f4(1.f), bt = beta ? ld4(beta + c4 * 4) : f4(0.f);
        return State{m, r, gm, bt, mul4(gm, r), mul4(ld4(sg + seg * C + c4 * 4), f4(inv_n)),
                     mul4(ld4(sgx + seg * C + c4 * 4), f4(inv_n))};
    }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        float4 gi = ld4(g + i);
        const float4 xh = mul4(sub4(ld4(x + i), st.mean), st.rstd);
        if (slope != 1.f) gi = mask4(gi, fma4(xh, st.gamma, st.beta), slope);
        st4(dx + i, mul4(st.coef, sub4(sub4(gi, st.a), mul4(xh, st.b))));
    }
};
// the same map with dx = (the BatchNorm backward) + addend[R, C] (its own type: the plain map's code is untouched -- a
// runtime `if (addend)` in the shared struct cost every BatchNorm backward 40 %)
struct NormBwdApplyAddOp4 {
    NormBwdApplyOp4 base; const float* addend;
    using State = NormBwdApplyOp4::State;
    __device__ State init(int c4, int64_t seg) const { return base.init(c4, seg); }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * base.C + c4 * 4;
        float4 gi = ld4(base.g + i);
        const float4 xh = mul4(sub4(ld4(base.x + i), st.mean), st.rstd);
        const float4 ad = ld4(addend + i);
        if (base.slope != 1.f) gi = mask4(gi, fma4(xh, st.gamma, st.beta), base.slope);
        st4(base.dx + i, add4(mul4(st.coef, sub4(sub4(gi, st.a), mul4(xh, st.b))), ad));
    }
};
struct DblBwdApplyOp4 {
    const float* g; const float* u; const float* x; int C; float inv_n; const float* mean; const float* rstd;
    const float* gamma; const float* sums; float* gg; float* gx;
    struct State { float4 mean, gm_r1, gm_r3, Su_n, Sg_n, Sux_n, Sgx_r3n, Sux_r3n, allsub_r3n; };
    __device__ State init(int c4, int64_t) const {
        const float4 r1 = ld4(rstd + c4 * 4);
        const float4 r2 = mul4(r1, r1), r3 = mul4(r2, r1);
        const float4 gm = gamma ? ld4(gamma + c4 * 4) : f4(1.f);
        const float4 Sg = ld4(sums + c4 * 4), Su = ld4(sums + C + c4 * 4), Sgx = ld4(sums + 2 * C + c4 * 4),
                     Sux = ld4(sums + 3 * C + c4 * 4), Sgu = ld4(sums + 4 * C + c4 * 4);
        const float4 n = f4(inv_n);
        // all_sub = Su*Sg/n - Sgu + 3 r^2 Sgx Sux / n
        const float4 all_sub = add4(sub4(mul4(mul4(Su, Sg), n), Sgu), mul4(mul4(f4(3.f), r2), mul4(mul4(Sgx, Sux), n)));
        return State{ld4(mean + c4 * 4), mul4(gm, r1), mul4(gm, r3), mul4(Su, n), mul4(Sg, n), mul4(Sux, n),
                     mul4(mul4(Sgx, r3), n), mul4(mul4(Sux, r3), n), mul4(mul4(all_sub, r3), n)};
    }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        const float4 gi = ld4(g + i), ui = ld4(u + i), xm = sub4(ld4(x + i), st.mean);
        // gg = gm r1 (u - Su/n) - gm r3 xm Sux/n
        st4(gg + i, sub4(mul4(st.gm_r1, sub4(ui, st.Su_n)), mul4(mul4(st.gm_r3, xm), st.Sux_n)));
        // gx = gm/r3-scaled terms: gm * (xm r3 all_sub/n + Sux r3/n (Sg/n - g) + Sgx r3/n (Su/n - u)); gm folded via gm_r1/r1
        const float4 t0 = mul4(xm, st.allsub_r3n);
        const float4 t1 = mul4(st.Sux_r3n, sub4(st.Sg_n, gi));
        const float4 t2 = mul4(st.Sgx_r3n, sub4(st.Su_n, ui));
        const float4 gmv = gamma ? ld4(gamma + c4 * 4) : f4(1.f);
        st4(gx + i, mul4(gmv, add4(add4(t0, t1), t2)));
    }
};
// Double backward of train-mode BatchNorm FUSED with LeakyReLU: the activation mask m (slope where the
// pre-activation is <= 0) is piecewise constant in x, so the closed form of plain BatchNorm applies to
// g' = m * g, and the cotangent reaching g is m * gg'.  The mask is recomputed from x on the fly.
struct DblBwdActOp4 {
    static constexpr int kStreams = 3;      // input tensors read per row (sets the unroll depth)
    const float* g; const float* u; const float* x; float slope; int C; const float* mean; const float* rstd;
    const float* gamma; const float* beta;
    struct State { float4 mean, rstd, gamma, beta; };
    __device__ State init(int c4, int64_t) const {
        return State{ld4(mean + c4 * 4), ld4(rstd + c4 * 4), gamma ? ld4(gamma + c4 * 4) : f4(1.f),
                     beta ? ld4(beta + c4 * 4) : f4(0.f)};
    }
    __device__ void accum(const State& st, int64_t r, int c4, float4* acc) const {
        const int64_t i = r * C + c4 * 4;
        const float4 ui = ld4(u + i), xm = sub4(ld4(x + i), st.mean);
        const float4 gi = mask4(ld4(g + i), fma4(mul4(xm, st.rstd), st.gamma, st.beta), slope);
        acc[0] = add4(acc[0], gi); acc[1] = add4(acc[1], ui); acc[2] = fma4(gi, xm, acc[2]);
        acc[3] = fma4(ui, xm, acc[3]); acc[4] = fma4(gi, ui, acc[4]);
    }
};
struct DblBwdActApplyOp4 {
    DblBwdApplyOp4 base; float slope; const float* beta;
    struct State { DblBwdApplyOp4::State b; float4 rstd, gamma, beta; };
    __device__ State init(int c4, int64_t seg) const {
        return State{base.init(c4, seg), ld4(base.rstd + c4 * 4), base.gamma ? ld4(base.gamma + c4 * 4) : f4(1.f),
                     beta ? ld4(beta + c4 * 4) : f4(0.f)};
    }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * base.C + c4 * 4;
        const float4 ui = ld4(base.u + i), xm = sub4(ld4(base.x + i), st.b.mean);
        const float4 pre = fma4(mul4(xm, st.rstd), st.gamma, st.beta);
        const float4 gi = mask4(ld4(base.g + i), pre, slope);
        const float4 ggp = sub4(mul4(st.b.gm_r1, sub4(ui, st.b.Su_n)), mul4(mul4(st.b.gm_r3, xm), st.b.Sux_n));
        st4(base.gg + i, mask4(ggp, pre, slope));
        const float4 t0 = mul4(xm, st.b.allsub_r3n);
        const float4 t1 = mul4(st.b.Sux_r3n, sub4(st.b.Sg_n, gi));
        const float4 t2 = mul4(st.b.Sgx_r3n, sub4(st.b.Su_n, ui));
        st4(base.gx + i, mul4(st.gamma, add4(add4(t0, t1), t2)));
    }
};
struct AdainApplyOp4 {
    const float* x; const float* s; int C; const float* mean; const float* rstd; float* out;
    struct State { float4 mean, rstd; };
    __device__ State init(int c4, int64_t seg) const { return State{ld4(mean + seg * C + c4 * 4), ld4(rstd + seg * C + c4 * 4)}; }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        const float4 xh = mul4(sub4(ld4(x + i), st.mean), st.rstd);
        st4(out + i, fma4(ld4(s + r * 2 * C + c4 * 4), xh, ld4(s + r * 2 * C + C + c4 * 4)));
    }
};
struct AdainBwdOp4 {
    const float* g; const float* x; const float* s; int C; const float* mean; const float* rstd; float* ds; float* gxh;
    struct State { float4 mean, rstd; };
    __device__ State init(int c4, int64_t seg) const { return State{ld4(mean + seg * C + c4 * 4), ld4(rstd + seg * C + c4 * 4)}; }
    __device__ void apply(const State& st, int64_t r, int c4) const {
        const int64_t i = r * C + c4 * 4;
        const float4 xh = mul4(sub4(ld4(x + i), st.mean), st.rstd);
        const float4 gi = ld4(g + i);
        if (ds) { st4(ds + r * 2 * C + c4 * 4, mul4(gi, xh)); st4(ds + r * 2 * C + C + c4 * 4, gi); }
        if (gxh) st4(gxh + i, mul4(gi, ld4(s + r * 2 * C + c4 * 4)));
    }
};

// ---------------------------------------------------------------------------------------
// normalisation maps
// ---------------------------------------------------------------------------------------
__global__ void norm_apply_kernel(const float* __restrict__ x, int64_t R, int C, int64_t seg_rows,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                  float* __restrict__ y) {
    const int64_t total = R * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const int64_t sc = (r / seg_rows) * C + c;
        float v = (__ldg(x + i) - __ldg(mean + sc)) * __ldg(rstd + sc);
        v = fmaf(v, gamma ? __ldg(gamma + c) : 1.f, beta ? __ldg(beta + c) : 0.f);
        y[i] = slope == 1.f ? v : lrelu_f(v, slope);
    }
}

// dx = gamma * rstd * (g' - sg/n - xhat * sgx/n)
__global__ void norm_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, float slope,
                                      int64_t R, int C, int64_t seg_rows, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ sg,
                                      const float* __restrict__ sgx, float* __restrict__ dx) {
    const int64_t total = R * C;
    const float inv_n = 1.f / (float)seg_rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const int64_t sc = (r / seg_rows) * C + c;
        float gi = __ldg(g + i);
        const float rs = __ldg(rstd + sc);
        const float xh = (__ldg(x + i) - __ldg(mean + sc)) * rs;
        const float gm = gamma ? __ldg(gamma + c) : 1.f;
        if (slope != 1.f && !(fmaf(xh, gm, beta ? __ldg(beta + c) : 0.f) > 0.f)) gi *= slope;
        dx[i] = gm * rs * (gi - __ldg(sg + sc) * inv_n - xh * __ldg(sgx + sc) * inv_n);
    }
}

// Double backward of train-mode BN w.r.t. (g, x, gamma) for cotangent u on dx.
// sums: [Sg, Su, Sgx, Sux, Sgu] with xm = x - mean (not normalised).
__global__ void bn_dbl_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ u,
                                        const float* __restrict__ x, int64_t R, int C,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ sums,
                                        float* __restrict__ gg, float* __restrict__ gx) {
    const int64_t total = R * C;
    const float inv_n = 1.f / (float)R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float r1 = __ldg(rstd + c);
        const float r2 = r1 * r1, r3 = r2 * r1;
        const float gm = gamma ? __ldg(gamma + c) : 1.f;
        const float Sg = __ldg(sums + c), Su = __ldg(sums + C + c), Sgx = __ldg(sums + 2 * C + c),
                    Sux = __ldg(sums + 3 * C + c), Sgu = __ldg(sums + 4 * C + c);
        const float gi = __ldg(g + i), ui = __ldg(u + i), xm = __ldg(x + i) - __ldg(mean + c);
        // d/dg: the (linear, self-adjoint) first-backward operator applied to u
        gg[i] = gm * r1 * (ui - Su * inv_n) - gm * r3 * xm * Sux * inv_n;
        // d/dx
        const float all_sub = Su * Sg * inv_n - Sgu + 3.f * r2 * Sgx * Sux * inv_n;
        const float t0 = xm * r3 * all_sub * inv_n;
        const float t1 = Sux * r3 * inv_n * (Sg * inv_n - gi);
        const float t2 = Sgx * r3 * inv_n * (Su * inv_n - ui);
        gx[i] = gm * (t0 + t1 + t2);
    }
}
__global__ void bn_dbl_bwd_gamma_kernel(int C, int64_t R, const float* __restrict__ rstd,
                                        const float* __restrict__ sums, float* __restrict__ ggamma) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float inv_n = 1.f / (float)R;
    const float r1 = rstd[c], r3 = r1 * r1 * r1;
    const float Sg = sums[c], Su = sums[C + c], Sgx = sums[2 * C + c], Sux = sums[3 * C + c], Sgu = sums[4 * C + c];
    ggamma[c] = r1 * Sgu - r1 * Su * Sg * inv_n - r3 * Sux * Sgx * inv_n;
}

__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ var, int C,
                                         float unbias, float momentum, float* __restrict__ rm,
                                         float* __restrict__ rv, int64_t* count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && count) *count += 1;
    if (c >= C) return;
    rm[c] = (1.f - momentum) * rm[c] + momentum * mean[c];
    rv[c] = (1.f - momentum) * rv[c] + momentum * (var[c] * unbias);
}

// ---------------------------------------------------------------------------------------
// max over the points of a segment (first arg max), scatter / gather by arg
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segmax_kernel(const float* __restrict__ x, int C, int64_t seg_rows, float* __restrict__ out,
              int32_t* __restrict__ arg) {
    __shared__ float bv[8][33];
    __shared__ int bi[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const int64_t seg = blockIdx.y;
    float best = -FLT_MAX;
    int besti = 0x7fffffff;
    if (c < C) {
        const float* base = x + seg * seg_rows * C + c;
#pragma unroll 4
        for (int64_t r = ty; r < seg_rows; r += 8) {
            const float v = __ldg(base + r * C);
            if (v > best || besti == 0x7fffffff) { best = v; besti = (int)r; }
        }
    }
    bv[ty][tx] = best; bi[ty][tx] = besti;
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int t = 1; t < 8; ++t) {
            const float v = bv[t][tx]; const int j = bi[t][tx];
            if (j != 0x7fffffff && (besti == 0x7fffffff || v > best || (v == best && j < besti))) { best = v; besti = j; }
        }
        out[seg * C + c] = best;
        if (arg) arg[seg * C + c] = besti;
    }
}

// float4 variant (C % 4 == 0, 16-byte aligned x): TX column groups of 4 channels x (256 / TX) row lanes per block,
// so every load is 16 bytes and a warp covers up to 512 contiguous bytes of a row.  Same result as segmax_kernel
// (max, first arg max) -- the combine below orders by (value desc, row asc).
template <int TX>
__global__ void __launch_bounds__(256)
segmax4_kernel(const float* __restrict__ x, int C4, int64_t seg_rows, float* __restrict__ out,
               int32_t* __restrict__ arg) {
    constexpr int TY = 256 / TX;
    __shared__ float4 bv[TY][TX];
    __shared__ int4 bi[TY][TX];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c4 = blockIdx.x * TX + tx;
    const int64_t seg = blockIdx.y;
    float best[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    int besti[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    if (c4 < C4) {
        const float4* base = reinterpret_cast<const float4*>(x + seg * seg_rows * (int64_t)C4 * 4) + c4;
#pragma unroll 4
        for (int64_t r = ty; r < seg_rows; r += TY) {
            const float4 v4 = __ldg(base + r * C4);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (v[q] > best[q] || besti[q] == 0x7fffffff) { best[q] = v[q]; besti[q] = (int)r; }
        }
    }
    bv[ty][tx] = make_float4(best[0], best[1], best[2], best[3]);
    bi[ty][tx] = make_int4(besti[0], besti[1], besti[2], besti[3]);
    __syncthreads();
    if (ty == 0 && c4 < C4) {
#pragma unroll 1
        for (int t = 1; t < TY; ++t) {
            const float4 v4 = bv[t][tx];
            const int4 j4 = bi[t][tx];
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            const int j[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j[q] != 0x7fffffff && (besti[q] == 0x7fffffff || v[q] > best[q] || (v[q] == best[q] && j[q] < besti[q]))) {
                    best[q] = v[q]; besti[q] = j[q];
                }
        }
        *reinterpret_cast<float4*>(out + (seg * C4 + c4) * 4) = make_float4(best[0], best[1], best[2], best[3]);
        if (arg) *reinterpret_cast<int4*>(arg + (seg * C4 + c4) * 4) = make_int4(besti[0], besti[1], besti[2], besti[3]);
    }
}

__global__ void segmax_scatter_kernel(const float* __restrict__ g, const int32_t* __restrict__ arg, int64_t nseg,
                                      int C, int64_t seg_rows, float* __restrict__ dx) {
    const int64_t total = nseg * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t seg = i / C;
        const int c = (int)(i - seg * C);
        dx[(seg * seg_rows + arg[i]) * C + c] = g[i];
    }
}
__global__ void segmax_gather_kernel(const float* __restrict__ x, const int32_t* __restrict__ arg, int64_t nseg,
                                     int C, int64_t seg_rows, float* __restrict__ out) {
    const int64_t total = nseg * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t seg = i / C;
        const int c = (int)(i - seg * C);
        out[i] = x[(seg * seg_rows + arg[i]) * C + c];
    }
}

// ---------------------------------------------------------------------------------------
// neighbour-axis ops on [P, k, C]
// ---------------------------------------------------------------------------------------
__global__ void softmax_k_kernel(const float* __restrict__ x, int64_t P, int k, int C, float* __restrict__ y) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const float* xp = x + p * k * C + c;
        float* yp = y + p * k * C + c;
        float m = -FLT_MAX;
        for (int r = 0; r < k; ++r) m = fmaxf(m, __ldg(xp + (int64_t)r * C));
        float s = 0.f;
        for (int r = 0; r < k; ++r) s += expf(__ldg(xp + (int64_t)r * C) - m);
        const float inv = 1.f / s;
        for (int r = 0; r < k; ++r) yp[(int64_t)r * C] = expf(__ldg(xp + (int64_t)r * C) - m) * inv;
    }
}
__global__ void softmax_k_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t P, int k,
                                     int C, float* __restrict__ dx) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t o = p * k * C + c;
        float s = 0.f;
        for (int r = 0; r < k; ++r) s = fmaf(__ldg(g + o + (int64_t)r * C), __ldg(y + o + (int64_t)r * C), s);
        for (int r = 0; r < k; ++r) {
            const float yy = __ldg(y + o + (int64_t)r * C);
            dx[o + (int64_t)r * C] = yy * (__ldg(g + o + (int64_t)r * C) - s);
        }
    }
}
// prod = y * softmax_k(x) (EdgeBlock: Generator.py:79,82) in one pass; w = softmax is kept for the backward
__global__ void softmax_mul_k_kernel(const float* __restrict__ x, const float* __restrict__ yv, int64_t P, int k, int C,
                                     float* __restrict__ w, float* __restrict__ prod) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t o = p * k * C + c;
        float m = -FLT_MAX;
        for (int r = 0; r < k; ++r) m = fmaxf(m, __ldg(x + o + (int64_t)r * C));
        float s = 0.f;
        for (int r = 0; r < k; ++r) s += expf(__ldg(x + o + (int64_t)r * C) - m);
        const float inv = 1.f / s;
        for (int r = 0; r < k; ++r) {
            const float wr = expf(__ldg(x + o + (int64_t)r * C) - m) * inv;
            w[o + (int64_t)r * C] = wr;
            prod[o + (int64_t)r * C] = wr * __ldg(yv + o + (int64_t)r * C);
        }
    }
}
// backward of prod = y * w, w = softmax_k(x): dy = g w, dx = w (g y - sum_r g y w)
// EdgeBlock attention with both train-mode BatchNorm + LeakyReLU applications folded into the loads
// (Generator.py:78-82: w = softmax_k(lrelu(bn(xw))), prod = lrelu(bn(xy)) * w).  Same fp32 operation order
// as norm_apply followed by softmax_mul_k, so the results are bit-identical to the unfused chain; the two
// normalised [E, C] tensors are never written.  k <= KMAXR values per thread stay in registers.
constexpr int KMAXR = 16;
struct BnCol { const float* mean; const float* rstd; const float* gamma; const float* beta; };
__device__ __forceinline__ float bn_act(float x, float m, float r, float g, float b, float slope) {
    return lrelu_f(fmaf((x - m) * r, g, b), slope);
}
template <int KM>
__global__ void __launch_bounds__(256)
bn_softmax_mul_k_kernel(const float* __restrict__ xw, const float* __restrict__ xy, int64_t P, int k, int C, BnCol bw,
                        BnCol by, float slope, float* __restrict__ w, float* __restrict__ prod) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t o = p * k * C + c;
        // all 2k loads of this (point, channel) are issued before anything depends on them
        float a[KM], b[KM];
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) { a[r] = __ldg(xw + o + (int64_t)r * C); b[r] = __ldg(xy + o + (int64_t)r * C); }
        const float mw = __ldg(bw.mean + c), rw = __ldg(bw.rstd + c), gw = __ldg(bw.gamma + c), bbw = __ldg(bw.beta + c);
        const float my = __ldg(by.mean + c), ry = __ldg(by.rstd + c), gy = __ldg(by.gamma + c), bby = __ldg(by.beta + c);
        float m = -FLT_MAX;
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) { a[r] = bn_act(a[r], mw, rw, gw, bbw, slope); m = fmaxf(m, a[r]); }
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) { a[r] = expf(a[r] - m); s += a[r]; }
        const float inv = 1.f / s;
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const float wr = a[r] * inv;
                if (w) w[o + (int64_t)r * C] = wr;               // the softmax weights are only kept for a backward pass
                prod[o + (int64_t)r * C] = wr * bn_act(b[r], my, ry, gy, bby, slope);
            }
    }
}
// backward of the fused forward up to the two activated tensors: dwa = d/d lrelu(bn(xw)), dya = d/d lrelu(bn(xy));
// lrelu(bn(xy)) is recomputed from xy.
template <int KM>
__global__ void __launch_bounds__(256)
bn_softmax_mul_k_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xy, const float* __restrict__ w,
                            int64_t P, int k, int C, BnCol by, float slope, float* __restrict__ dwa,
                            float* __restrict__ dya) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t o = p * k * C + c;
        // all 3k loads of this (point, channel) are issued before anything depends on them (a store between the
        // loads of consecutive neighbours serialises them: 3 loads in flight per thread, ~3.7 TB/s)
        float gv[KM], wv[KM], yv[KM];
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const int64_t a = o + (int64_t)r * C;
                gv[r] = __ldg(g + a);
                wv[r] = __ldg(w + a);
                yv[r] = __ldg(xy + a);
            }
        const float my = __ldg(by.mean + c), ry = __ldg(by.rstd + c), gy = __ldg(by.gamma + c), bby = __ldg(by.beta + c);
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                yv[r] = gv[r] * bn_act(yv[r], my, ry, gy, bby, slope);      // g * lrelu(bn(xy))
                s = fmaf(yv[r], wv[r], s);
            }
#pragma unroll
        for (int r = 0; r < KM; ++r)
            if (r < k) {
                const int64_t a = o + (int64_t)r * C;
                if (dya) dya[a] = gv[r] * wv[r];
                if (dwa) dwa[a] = wv[r] * (yv[r] - s);
            }
    }
}
__global__ void softmax_mul_k_bwd_kernel(const float* __restrict__ g, const float* __restrict__ yv,
                                         const float* __restrict__ w, int64_t P, int k, int C, float* __restrict__ dx,
                                         float* __restrict__ dy) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const int64_t o = p * k * C + c;
        float s = 0.f;
        for (int r = 0; r < k; ++r) {
            const int64_t a = o + (int64_t)r * C;
            s = fmaf(__ldg(g + a) * __ldg(yv + a), __ldg(w + a), s);
        }
        for (int r = 0; r < k; ++r) {
            const int64_t a = o + (int64_t)r * C;
            const float gr = __ldg(g + a), wr = __ldg(w + a);
            if (dx) dx[a] = wr * (gr * __ldg(yv + a) - s);
            if (dy) dy[a] = gr * wr;
        }
    }
}
__global__ void kmax_kernel(const float* __restrict__ x, int64_t P, int k, int C, float* __restrict__ out,
                            int32_t* __restrict__ arg) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const float* xp = x + p * k * C + c;
        float m = __ldg(xp);
        int a = 0;
        for (int r = 1; r < k; ++r) {
            const float v = __ldg(xp + (int64_t)r * C);
            if (v > m) { m = v; a = r; }
        }
        out[i] = m;
        if (arg) arg[i] = a;
    }
}
__global__ void kmax_scatter_kernel(const float* __restrict__ g, const int32_t* __restrict__ arg, int64_t P, int k,
                                    int C, float* __restrict__ dx) {
    const int64_t total = P * k * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int r = (int)((i / C) % k);
        const int64_t p = i / ((int64_t)C * k);
        dx[i] = (arg[p * C + c] == r) ? g[p * C + c] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// edge aggregation through per-point projections
// ---------------------------------------------------------------------------------------
template <bool VEC>
__global__ void edge_combine_kernel(const float* __restrict__ pc, const float* __restrict__ pn,
                                    const int32_t* __restrict__ idx, const float* __restrict__ bias, int64_t P, int N,
                                    int k, int C, float* __restrict__ out) {
    // one thread per (point, channel group): the centre terms pc[p] - pn[p] + bias are loaded once and reused
    // for the k neighbours (k + 2 row loads per k outputs instead of 3k)
    constexpr int W = VEC ? 4 : 1;
    const int Cw = C / W;
    const int64_t total = P * Cw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cw = (int)(i % Cw);
        const int64_t p = i / Cw;
        const int64_t base = (p / N) * N;
        const int32_t* ip = idx + p * k;
        if (VEC) {
            // same operation order as the reference expression (pn[j] - pn[p]) + pc[p] + bias
            const float4 b = __ldg(reinterpret_cast<const float4*>(pn + p * C) + cw);
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f), bb = q;
            if (pc) q = __ldg(reinterpret_cast<const float4*>(pc + p * C) + cw);
            if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias) + cw);
            float4* o = reinterpret_cast<float4*>(out) + p * k * Cw + cw;
#pragma unroll 5
            for (int r = 0; r < k; ++r) {
                const int64_t j = base + __ldg(ip + r);
                const float4 a = __ldg(reinterpret_cast<const float4*>(pn + j * C) + cw);
                float4 v = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
                if (pc) { v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
                if (bias) { v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
                o[(int64_t)r * Cw] = v;
            }
        } else {
            const float b = __ldg(pn + p * C + cw);
            const float q = pc ? __ldg(pc + p * C + cw) : 0.f;
            const float bb = bias ? __ldg(bias + cw) : 0.f;
            for (int r = 0; r < k; ++r) {
                const int64_t j = base + __ldg(ip + r);
                float v = __ldg(pn + j * C + cw) - b;
                if (pc) v += q;
                if (bias) v += bb;
                out[(p * k + r) * C + cw] = v;
            }
        }
    }
}

template <bool VEC>
__global__ void edge_combine_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx, int64_t P,
                                        int N, int k, int C, float* __restrict__ dpc, float* __restrict__ dpn) {
    constexpr int W = VEC ? 4 : 1;
    const int Cw = C / W;
    const int64_t total = P * Cw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cw = (int)(i % Cw);
        const int64_t p = i / Cw;
        const int64_t base = (p / N) * N;
        if (VEC) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < k; ++r) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(g + (p * k + r) * C) + cw);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                const int64_t j = base + __ldg(idx + p * k + r);
                atomicAdd(reinterpret_cast<float4*>(dpn + j * C) + cw, v);
            }
            if (dpc) reinterpret_cast<float4*>(dpc + p * C)[cw] = s;
            atomicAdd(reinterpret_cast<float4*>(dpn + p * C) + cw, make_float4(-s.x, -s.y, -s.z, -s.w));
        } else {
            float s = 0.f;
            for (int r = 0; r < k; ++r) {
                const float v = __ldg(g + (p * k + r) * C + cw);
                s += v;
                const int64_t j = base + __ldg(idx + p * k + r);
                atomicAdd(dpn + j * C + cw, v);
            }
            if (dpc) dpc[p * C + cw] = s;
            atomicAdd(dpn + p * C + cw, -s);
        }
    }
}

// ---------------------------------------------------------------------------------------
// AdaIN
// ---------------------------------------------------------------------------------------
__global__ void adain_apply_kernel(const float* __restrict__ x, const float* __restrict__ s, int64_t R, int C,
                                   int64_t seg_rows, const float* __restrict__ mean, const float* __restrict__ rstd,
                                   float* __restrict__ out) {
    const int64_t total = R * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const int64_t sc = (r / seg_rows) * C + c;
        const float xh = (__ldg(x + i) - __ldg(mean + sc)) * __ldg(rstd + sc);
        out[i] = fmaf(__ldg(s + r * 2 * C + c), xh, __ldg(s + r * 2 * C + C + c));
    }
}
__global__ void adain_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                 const float* __restrict__ s, int64_t R, int C, int64_t seg_rows,
                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                 float* __restrict__ ds, float* __restrict__ gxh) {
    const int64_t total = R * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const int64_t sc = (r / seg_rows) * C + c;
        const float xh = (__ldg(x + i) - __ldg(mean + sc)) * __ldg(rstd + sc);
        const float gi = __ldg(g + i);
        if (ds) { ds[r * 2 * C + c] = gi * xh; ds[r * 2 * C + C + c] = gi; }
        if (gxh) gxh[i] = gi * __ldg(s + r * 2 * C + c);
    }
}

// ---------------------------------------------------------------------------------------
// gradient penalty reductions
// ---------------------------------------------------------------------------------------
__device__ float block_sum(float v) {
    __shared__ float sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;   // valid in warp 0
}
__global__ void gp_norm_kernel(const float* __restrict__ g, int64_t D, float* __restrict__ norms) {
    const float* gb = g + blockIdx.x * D;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < D; i += blockDim.x) { const float v = __ldg(gb + i); s = fmaf(v, v, s); }
    s = block_sum(s);
    if (threadIdx.x == 0) norms[blockIdx.x] = sqrtf(s);
}
__global__ void gp_penalty_kernel(const float* __restrict__ norms, int B, float gamma, float lambda,
                                  float* __restrict__ penalty) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { const float t = (norms[i] - gamma) / gamma; s = fmaf(t, t, s); }
    s = block_sum(s);
    if (threadIdx.x == 0) *penalty = lambda * s / (float)B;
}
__global__ void gp_penalty_bwd_kernel(const float* __restrict__ g, const float* __restrict__ norms,
                                      const float* __restrict__ gout, int B, int64_t D, float gamma, float lambda,
                                      float* __restrict__ dg) {
    const int64_t total = (int64_t)B * D;
    const float go = __ldg(gout);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / D);
        const float nrm = __ldg(norms + b);
        const float coef = go * lambda * 2.f * (nrm - gamma) / (gamma * gamma * (float)B * nrm);
        dg[i] = coef * __ldg(g + i);
    }
}
__global__ void mean_kernel(const float* __restrict__ x, int64_t n, float scale, int accumulate,
                            float* __restrict__ out) {
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += __ldg(x + i);
    s = block_sum(s);
    if (threadIdx.x == 0) {
        const float v = scale * s / (float)n;
        *out = accumulate ? *out + v : v;
    }
}

}  // namespace

// =========================================================================================
extern "C" size_t spgan_colreduce_workspace(int64_t R, int C, int64_t seg_rows, int nvals) {
    if (R <= 0 || C <= 0 || seg_rows <= 0 || nvals <= 0 || R % seg_rows != 0) return 0;
    const ChunkPlan p = plan_chunks(R, C, seg_rows);
    size_t n = (size_t)p.nseg * p.chunks;
    if (C % 4 == 0) {
        const fastnorm::Plan q = fastnorm::make_plan(R, C / 4, fastnorm::pick_tx(C / 4), seg_rows, kReduceTargetBlocks);
        const size_t m = (size_t)q.nseg * q.chunks;
        if (m > n) n = m;
    }
    return n * nvals * C * sizeof(float);
}

extern "C" int spgan_colsum(const float* x, int64_t R, int C, int64_t seg_rows, float* out, void* ws,
                            spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && out && R >= 0 && C >= 1);
    return run_colreduce<1>(R, C, seg_rows, ws, as_stream(s), SumOp{x, C}, SumOp4{x, C}, al16(x), Store1Fin{out, C});
}
extern "C" int spgan_coldot(const float* x, const float* y, int64_t R, int C, int64_t seg_rows, float* out,
                            void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && y && out && R >= 0 && C >= 1);
    return run_colreduce<1>(R, C, seg_rows, ws, as_stream(s), DotOp{x, y, C}, DotOp4{x, y, C}, al16(x) && al16(y),
                            Store1Fin{out, C});
}
extern "C" int spgan_colstats(const float* x, int64_t R, int C, int64_t seg_rows, float eps, float* mean,
                              float* rstd, float* var, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && mean && rstd && R >= 1 && C >= 1);
    return run_colreduce<2>(R, C, seg_rows, ws, as_stream(s), StatsOp{x, C, seg_rows}, StatsOp4{x, C, seg_rows}, al16(x),
                            StatsFin{x, C, seg_rows, eps, mean, rstd, var, 0.f, 1.f, nullptr, nullptr, nullptr});
}
extern "C" int spgan_norm_apply(const float* x, int64_t R, int C, int64_t seg_rows, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, float slope, float* y,
                                spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && mean && rstd && y && R >= 0 && C >= 1 && seg_rows >= 1);
    if (R == 0) return SPGAN_OK;
    if (C % 4 == 0 && R % seg_rows == 0 && al16(x) && al16(y) && al16(mean) && al16(rstd) && (!gamma || al16(gamma)) &&
        (!beta || al16(beta)))
        return fastnorm::run_map(R, C, seg_rows, as_stream(s), NormApplyOp4{x, C, mean, rstd, gamma, beta, slope, y});
    norm_apply_kernel<<<ew_grid(R * C, 256, 16), 256, 0, as_stream(s)>>>(x, R, C, seg_rows, mean, rstd, gamma, beta,
                                                                         slope, y);
    return spgan_launch_status();
}
extern "C" int spgan_bn_update_running(const float* mean, const float* var, int C, int64_t R, float momentum,
                                       float* rm, float* rv, int64_t* count, spgan_stream_t s) {
    SPGAN_CHECK_ARG(mean && var && rm && rv && C >= 1 && R >= 1);
    const float unbias = R > 1 ? (float)((double)R / (double)(R - 1)) : 1.f;
    bn_update_running_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(mean, var, C, unbias, momentum, rm, rv, count);
    return spgan_launch_status();
}
extern "C" int spgan_colstats_bn(const float* x, int64_t R, int C, float eps, float* mean, float* rstd, float* var,
                                 float momentum, float* rm, float* rv, int64_t* count, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && mean && rstd && var && rm && rv && R >= 1 && C >= 1);
    const float unbias = R > 1 ? (float)((double)R / (double)(R - 1)) : 1.f;
    return run_colreduce<2>(R, C, R, ws, as_stream(s), StatsOp{x, C, R}, StatsOp4{x, C, R}, al16(x),
                            StatsFin{x, C, R, eps, mean, rstd, var, momentum, unbias, rm, rv, count});
}

// Column partial sums of a GEMM epilogue (spgan_gemm_fused) -> batch statistics, the consumer's prologue tables and
// the running-statistics update; partials added in double along a fixed tree (deterministic).
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ ps, const float* __restrict__ pq, int rows, int C, double inv_n,
                   float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ var,
                   float* __restrict__ scale, float* __restrict__ shift, float momentum, float unbias,
                   float* rm, float* rv, int64_t* count) {
    // block = 32 columns x 32 row lanes; every lane adds its rows in order, the 32 lane sums are combined in order:
    // a fixed summation tree (deterministic), in double
    __shared__ double red[2][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double s1 = 0.0, s2 = 0.0;
    if (c < C)
        for (int r = ty; r < rows; r += 32) { s1 += (double)__ldg(ps + (int64_t)r * C + c); s2 += (double)__ldg(pq + (int64_t)r * C + c); }
    red[0][ty][tx] = s1;
    red[1][ty][tx] = s2;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    s1 = 0.0; s2 = 0.0;
    for (int t = 0; t < 32; ++t) { s1 += red[0][t][tx]; s2 += red[1][t][tx]; }
    const double m = s1 * inv_n;
    double v = s2 * inv_n - m * m;
    if (v < 0.0) v = 0.0;
    const float mf = (float)m, rs = (float)(1.0 / sqrt(v + (double)eps));
    mean[c] = mf; rstd[c] = rs; var[c] = (float)v;
    if (scale) {
        // y = ((x - mean) * rstd) * gamma + beta = x * scale + shift
        const float sc = rs * (gamma ? gamma[c] : 1.f);
        scale[c] = sc;
        shift[c] = fmaf(-mf, sc, beta ? beta[c] : 0.f);
    }
    if (rm) {
        rm[c] = (1.f - momentum) * rm[c] + momentum * mf;
        rv[c] = (1.f - momentum) * rv[c] + momentum * ((float)v * unbias);
        if (c == 0 && count) *count += 1;
    }
}
extern "C" int spgan_bn_finalize(const float* col_sum, const float* col_sqsum, int rows, int C, int64_t R, float eps,
                                 const float* gamma, const float* beta, float* mean, float* rstd, float* var,
                                 float* scale, float* shift, float momentum, float* rm, float* rv, int64_t* count,
                                 spgan_stream_t s) {
    SPGAN_CHECK_ARG(col_sum && col_sqsum && mean && rstd && var && rows >= 1 && C >= 1 && R >= 1);
    SPGAN_CHECK_ARG((scale == nullptr) == (shift == nullptr) && (rm == nullptr) == (rv == nullptr));
    const float unbias = R > 1 ? (float)((double)R / (double)(R - 1)) : 1.f;
    bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, as_stream(s)>>>(col_sum, col_sqsum, rows, C, 1.0 / (double)R, eps, gamma, beta,
                                                               mean, rstd, var, scale, shift, momentum, unbias, rm, rv, count);
    return spgan_launch_status();
}
// (mean, rstd, gamma, beta) -> the prologue tables of spgan_gemm_fused (eval-mode statistics, or a saved batch)
__global__ void bn_tables_kernel(const float* mean, const float* rstd, const float* gamma, const float* beta, int C,
                                 float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = rstd[c] * (gamma ? gamma[c] : 1.f);
    scale[c] = sc;
    shift[c] = fmaf(-mean[c], sc, beta ? beta[c] : 0.f);
}
extern "C" int spgan_bn_tables(const float* mean, const float* rstd, const float* gamma, const float* beta, int C,
                               float* scale, float* shift, spgan_stream_t s) {
    SPGAN_CHECK_ARG(mean && rstd && scale && shift && C >= 1);
    bn_tables_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(mean, rstd, gamma, beta, C, scale, shift);
    return spgan_launch_status();
}
extern "C" int spgan_norm_bwd_reduce_acc(const float* g, const float* x, float slope, int64_t R, int C, const float* mean,
                                         const float* rstd, const float* gamma, const float* beta, float* sg, float* sgx,
                                         float* acc_dbeta, float* acc_dgamma, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && mean && rstd && sg && sgx && R >= 1 && C >= 1);
    const bool vec = al16(g) && al16(x) && al16(mean) && al16(rstd) && (!gamma || al16(gamma)) && (!beta || al16(beta));
    return run_colreduce<2>(R, C, R, ws, as_stream(s), NormBwdOp{g, x, slope, C, mean, rstd, gamma, beta},
                            NormBwdOp4{g, x, slope, C, mean, rstd, gamma, beta}, vec,
                            Store2Fin{sg, sgx, C, acc_dbeta, acc_dgamma});
}
extern "C" int spgan_norm_bwd_reduce(const float* g, const float* x, float slope, int64_t R, int C,
                                     int64_t seg_rows, const float* mean, const float* rstd, const float* gamma,
                                     const float* beta, float* sg, float* sgx, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && mean && rstd && sg && sgx && R >= 1 && C >= 1);
    const bool vec = al16(g) && al16(x) && al16(mean) && al16(rstd) && (!gamma || al16(gamma)) && (!beta || al16(beta));
    return run_colreduce<2>(R, C, seg_rows, ws, as_stream(s), NormBwdOp{g, x, slope, C, mean, rstd, gamma, beta},
                            NormBwdOp4{g, x, slope, C, mean, rstd, gamma, beta}, vec, Store2Fin{sg, sgx, C, nullptr, nullptr});
}
extern "C" int spgan_norm_bwd_apply(const float* g, const float* x, float slope, int64_t R, int C, int64_t seg_rows,
                                    const float* mean, const float* rstd, const float* gamma, const float* beta,
                                    const float* sg, const float* sgx, float* dx, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && mean && rstd && sg && sgx && dx && R >= 1 && C >= 1 && seg_rows >= 1);
    if (C % 4 == 0 && R % seg_rows == 0 && al16(g) && al16(x) && al16(dx) && al16(mean) && al16(rstd) && al16(sg) &&
        al16(sgx) && (!gamma || al16(gamma)) && (!beta || al16(beta)))
        return fastnorm::run_map(R, C, seg_rows, as_stream(s),
                                 NormBwdApplyOp4{g, x, slope, C, 1.f / (float)seg_rows, mean, rstd, gamma, beta, sg, sgx, dx});
    norm_bwd_apply_kernel<<<ew_grid(R * C, 256, 16), 256, 0, as_stream(s)>>>(g, x, slope, R, C, seg_rows, mean, rstd,
                                                                             gamma, beta, sg, sgx, dx);
    return spgan_launch_status();
}
extern "C" int spgan_norm_bwd_apply_add(const float* g, const float* x, float slope, int64_t R, int C, const float* mean,
                                        const float* rstd, const float* gamma, const float* beta, const float* sg,
                                        const float* sgx, const float* addend, float* dx, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && mean && rstd && sg && sgx && addend && dx && R >= 1 && C >= 1);
    if (!(C % 4 == 0 && al16(g) && al16(x) && al16(dx) && al16(addend) && al16(mean) && al16(rstd) && al16(sg) && al16(sgx) &&
          (!gamma || al16(gamma)) && (!beta || al16(beta))))
        return SPGAN_E_UNSUPPORTED;
    return fastnorm::run_map(R, C, R, as_stream(s),
                             NormBwdApplyAddOp4{NormBwdApplyOp4{g, x, slope, C, 1.f / (float)R, mean, rstd, gamma, beta, sg, sgx, dx},
                                                addend});
}
extern "C" int spgan_bn_dbl_bwd_reduce(const float* g, const float* u, const float* x, int64_t R, int C,
                                       const float* mean, float* sums, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && u && x && mean && sums && R >= 1 && C >= 1);
    return run_colreduce<5>(R, C, R, ws, as_stream(s), DblBwdOp{g, u, x, C, mean}, DblBwdOp4{g, u, x, C, mean},
                            al16(g) && al16(u) && al16(x) && al16(mean), Store5Fin{sums, C});
}
extern "C" int spgan_bn_dbl_bwd_apply(const float* g, const float* u, const float* x, int64_t R, int C,
                                      const float* mean, const float* rstd, const float* gamma, const float* sums,
                                      float* gg, float* gx, float* ggamma, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && u && x && mean && rstd && sums && gg && gx && R >= 1 && C >= 1);
    if (C % 4 == 0 && al16(g) && al16(u) && al16(x) && al16(gg) && al16(gx) && al16(mean) && al16(rstd) && al16(sums) &&
        (!gamma || al16(gamma))) {
        int rc = fastnorm::run_map(R, C, R, as_stream(s),
                                   DblBwdApplyOp4{g, u, x, C, 1.f / (float)R, mean, rstd, gamma, sums, gg, gx});
        if (rc != SPGAN_OK) return rc;
    } else {
        bn_dbl_bwd_apply_kernel<<<ew_grid(R * C, 256, 16), 256, 0, as_stream(s)>>>(g, u, x, R, C, mean, rstd, gamma, sums,
                                                                                   gg, gx);
    }
    if (ggamma) bn_dbl_bwd_gamma_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(C, R, rstd, sums, ggamma);
    return spgan_launch_status();
}

extern "C" int spgan_bn_act_dbl_bwd_reduce(const float* g, const float* u, const float* x, float slope, int64_t R,
                                           int C, const float* mean, const float* rstd, const float* gamma,
                                           const float* beta, float* sums, void* ws, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && u && x && mean && rstd && sums && ws && R >= 1 && C >= 1);
    if (C % 4 != 0 || !(al16(g) && al16(u) && al16(x) && al16(mean) && al16(rstd) && (!gamma || al16(gamma)) &&
                        (!beta || al16(beta))))
        return SPGAN_E_UNSUPPORTED;
    return run_colreduce<5>(R, C, R, ws, as_stream(s), DblBwdOp{g, u, x, C, mean},
                            DblBwdActOp4{g, u, x, slope, C, mean, rstd, gamma, beta}, true, Store5Fin{sums, C});
}
extern "C" int spgan_bn_act_dbl_bwd_apply(const float* g, const float* u, const float* x, float slope, int64_t R,
                                          int C, const float* mean, const float* rstd, const float* gamma,
                                          const float* beta, const float* sums, float* gg, float* gx, float* ggamma,
                                          spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && u && x && mean && rstd && sums && gg && gx && R >= 1 && C >= 1);
    if (C % 4 != 0 || !(al16(g) && al16(u) && al16(x) && al16(gg) && al16(gx) && al16(mean) && al16(rstd) &&
                        al16(sums) && (!gamma || al16(gamma)) && (!beta || al16(beta))))
        return SPGAN_E_UNSUPPORTED;
    int rc = fastnorm::run_map(R, C, R, as_stream(s),
                               DblBwdActApplyOp4{DblBwdApplyOp4{g, u, x, C, 1.f / (float)R, mean, rstd, gamma, sums, gg, gx},
                                                 slope, beta});
    if (rc != SPGAN_OK) return rc;
    if (ggamma) bn_dbl_bwd_gamma_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(C, R, rstd, sums, ggamma);
    return spgan_launch_status();
}

extern "C" int spgan_segmax(const float* x, int64_t R, int C, int64_t seg_rows, float* out, int32_t* arg,
                            spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && out && R >= 0 && C >= 1 && seg_rows >= 1 && R % seg_rows == 0);
    if (R == 0) return SPGAN_OK;
    const int64_t nseg = R / seg_rows;
    if (nseg > 65535) return SPGAN_E_UNSUPPORTED;
    const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                        reinterpret_cast<uintptr_t>(arg)) & 15) == 0;
    if (C % 4 == 0 && al16) {
        const int C4 = C / 4;
        if (C4 >= 64) segmax4_kernel<32><<<dim3((C4 + 31) / 32, (unsigned)nseg), 256, 0, as_stream(s)>>>(x, C4, seg_rows, out, arg);
        else segmax4_kernel<8><<<dim3((C4 + 7) / 8, (unsigned)nseg), 256, 0, as_stream(s)>>>(x, C4, seg_rows, out, arg);
        return spgan_launch_status();
    }
    dim3 grid((C + 31) / 32, (unsigned)nseg), block(32, 8);
    segmax_kernel<<<grid, block, 0, as_stream(s)>>>(x, C, seg_rows, out, arg);
    return spgan_launch_status();
}
extern "C" int spgan_segmax_scatter(const float* g, const int32_t* arg, int64_t R, int C, int64_t seg_rows,
                                    float* dx, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && arg && dx && R >= 0 && C >= 1 && seg_rows >= 1 && R % seg_rows == 0);
    if (R == 0) return SPGAN_OK;
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)R * C * sizeof(float), as_stream(s));
    if (e != cudaSuccess) return (int)e;
    const int64_t nseg = R / seg_rows;
    segmax_scatter_kernel<<<ew_grid(nseg * C, 256), 256, 0, as_stream(s)>>>(g, arg, nseg, C, seg_rows, dx);
    return spgan_launch_status();
}
extern "C" int spgan_segmax_gather(const float* x, const int32_t* arg, int64_t R, int C, int64_t seg_rows,
                                   float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && arg && out && R >= 0 && C >= 1 && seg_rows >= 1 && R % seg_rows == 0);
    if (R == 0) return SPGAN_OK;
    const int64_t nseg = R / seg_rows;
    segmax_gather_kernel<<<ew_grid(nseg * C, 256), 256, 0, as_stream(s)>>>(x, arg, nseg, C, seg_rows, out);
    return spgan_launch_status();
}
extern "C" int spgan_softmax_k(const float* x, int64_t P, int k, int C, float* y, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && y && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    softmax_k_kernel<<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(x, P, k, C, y);
    return spgan_launch_status();
}
extern "C" int spgan_softmax_k_bwd(const float* g, const float* y, int64_t P, int k, int C, float* dx,
                                   spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && y && dx && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    softmax_k_bwd_kernel<<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(g, y, P, k, C, dx);
    return spgan_launch_status();
}
extern "C" int spgan_softmax_mul_k(const float* x, const float* y, int64_t P, int k, int C, float* w, float* prod,
                                   spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && y && w && prod && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    softmax_mul_k_kernel<<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(x, y, P, k, C, w, prod);
    return spgan_launch_status();
}
extern "C" int spgan_softmax_mul_k_bwd(const float* g, const float* y, const float* w, int64_t P, int k, int C,
                                       float* dx, float* dy, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && y && w && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    softmax_mul_k_bwd_kernel<<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(g, y, w, P, k, C, dx, dy);
    return spgan_launch_status();
}
extern "C" int spgan_bn_softmax_mul_k(const float* xw, const float* xy, int64_t P, int k, int C, const float* mean_w,
                                      const float* rstd_w, const float* gamma_w, const float* beta_w,
                                      const float* mean_y, const float* rstd_y, const float* gamma_y,
                                      const float* beta_y, float slope, float* w, float* prod, spgan_stream_t s) {
    SPGAN_CHECK_ARG(xw && xy && prod && mean_w && rstd_w && gamma_w && beta_w && mean_y && rstd_y && gamma_y && beta_y);
    SPGAN_CHECK_ARG(P >= 0 && k >= 1 && C >= 1);
    if (k > KMAXR) return SPGAN_E_UNSUPPORTED;
    if (P == 0) return SPGAN_OK;
    const BnCol cw{mean_w, rstd_w, gamma_w, beta_w}, cy{mean_y, rstd_y, gamma_y, beta_y};
    if (k <= 10)
        bn_softmax_mul_k_kernel<10><<<ew_grid(P * C, 256, 32), 256, 0, as_stream(s)>>>(xw, xy, P, k, C, cw, cy, slope, w, prod);
    else
        bn_softmax_mul_k_kernel<KMAXR><<<ew_grid(P * C, 256, 32), 256, 0, as_stream(s)>>>(xw, xy, P, k, C, cw, cy, slope, w, prod);
    return spgan_launch_status();
}
extern "C" int spgan_bn_softmax_mul_k_bwd(const float* g, const float* xy, const float* w, int64_t P, int k, int C,
                                          const float* mean_y, const float* rstd_y, const float* gamma_y,
                                          const float* beta_y, float slope, float* dwa, float* dya, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && xy && w && mean_y && rstd_y && gamma_y && beta_y && P >= 0 && k >= 1 && C >= 1);
    if (k > KMAXR) return SPGAN_E_UNSUPPORTED;
    if (P == 0) return SPGAN_OK;
    const BnCol cy{mean_y, rstd_y, gamma_y, beta_y};
    if (k <= 10)
        bn_softmax_mul_k_bwd_kernel<10><<<ew_grid(P * C, 256, 32), 256, 0, as_stream(s)>>>(g, xy, w, P, k, C, cy, slope, dwa, dya);
    else
        bn_softmax_mul_k_bwd_kernel<KMAXR><<<ew_grid(P * C, 256, 32), 256, 0, as_stream(s)>>>(g, xy, w, P, k, C, cy, slope, dwa, dya);
    return spgan_launch_status();
}
extern "C" int spgan_kmax(const float* x, int64_t P, int k, int C, float* out, int32_t* arg, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && out && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    kmax_kernel<<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(x, P, k, C, out, arg);
    return spgan_launch_status();
}
extern "C" int spgan_kmax_scatter(const float* g, const int32_t* arg, int64_t P, int k, int C, float* dx,
                                  spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && arg && dx && P >= 0 && k >= 1 && C >= 1);
    if (P == 0) return SPGAN_OK;
    kmax_scatter_kernel<<<ew_grid(P * k * C, 256, 16), 256, 0, as_stream(s)>>>(g, arg, P, k, C, dx);
    return spgan_launch_status();
}
extern "C" int spgan_edge_combine(const float* pc, const float* pn, const int32_t* idx, const float* bias, int64_t P,
                                  int N, int k, int C, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(pn && idx && out && P >= 0 && N >= 1 && k >= 1 && C >= 1 && P % N == 0);
    if (P == 0) return SPGAN_OK;
    const bool vec = (C % 4 == 0) && al16(pn) && al16(out) && (!pc || al16(pc)) && (!bias || al16(bias));
    if (vec) edge_combine_kernel<true><<<ew_grid(P * (C / 4), 256, 16), 256, 0, as_stream(s)>>>(pc, pn, idx, bias, P, N, k, C, out);
    else edge_combine_kernel<false><<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(pc, pn, idx, bias, P, N, k, C, out);
    return spgan_launch_status();
}
extern "C" int spgan_edge_combine_bwd(const float* g, const int32_t* idx, int64_t P, int N, int k, int C, float* dpc,
                                      float* dpn, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && idx && dpn && P >= 0 && N >= 1 && k >= 1 && C >= 1 && P % N == 0);
    if (P == 0) return SPGAN_OK;
    cudaError_t e = cudaMemsetAsync(dpn, 0, (size_t)P * C * sizeof(float), as_stream(s));
    if (e != cudaSuccess) return (int)e;
    const bool vec = (C % 4 == 0) && al16(g) && al16(dpn) && (!dpc || al16(dpc));
    if (vec) edge_combine_bwd_kernel<true><<<ew_grid(P * (C / 4), 256, 16), 256, 0, as_stream(s)>>>(g, idx, P, N, k, C, dpc, dpn);
    else edge_combine_bwd_kernel<false><<<ew_grid(P * C, 256, 16), 256, 0, as_stream(s)>>>(g, idx, P, N, k, C, dpc, dpn);
    return spgan_launch_status();
}
extern "C" int spgan_adain_apply(const float* x, const float* sv, int64_t R, int C, int64_t seg_rows,
                                 const float* mean, const float* rstd, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && sv && mean && rstd && out && R >= 0 && C >= 1 && seg_rows >= 1);
    if (R == 0) return SPGAN_OK;
    if (C % 4 == 0 && R % seg_rows == 0 && al16(x) && al16(sv) && al16(out) && al16(mean) && al16(rstd))
        return fastnorm::run_map(R, C, seg_rows, as_stream(s), AdainApplyOp4{x, sv, C, mean, rstd, out});
    adain_apply_kernel<<<ew_grid(R * C, 256, 16), 256, 0, as_stream(s)>>>(x, sv, R, C, seg_rows, mean, rstd, out);
    return spgan_launch_status();
}
extern "C" int spgan_adain_bwd(const float* g, const float* x, const float* sv, int64_t R, int C, int64_t seg_rows,
                               const float* mean, const float* rstd, float* ds, float* gxh, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && x && sv && mean && rstd && R >= 0 && C >= 1 && seg_rows >= 1);
    if (R == 0) return SPGAN_OK;
    if (C % 4 == 0 && R % seg_rows == 0 && al16(g) && al16(x) && al16(sv) && al16(mean) && al16(rstd) &&
        (!ds || al16(ds)) && (!gxh || al16(gxh)))
        return fastnorm::run_map(R, C, seg_rows, as_stream(s), AdainBwdOp4{g, x, sv, C, mean, rstd, ds, gxh});
    adain_bwd_kernel<<<ew_grid(R * C, 256, 16), 256, 0, as_stream(s)>>>(g, x, sv, R, C, seg_rows, mean, rstd, ds, gxh);
    return spgan_launch_status();
}
extern "C" int spgan_gp_penalty(const float* g, int B, int64_t D, float gamma, float lambda, float* norms,
                                float* penalty, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && norms && penalty && B >= 1 && D >= 1);
    gp_norm_kernel<<<B, 256, 0, as_stream(s)>>>(g, D, norms);
    gp_penalty_kernel<<<1, 256, 0, as_stream(s)>>>(norms, B, gamma, lambda, penalty);
    return spgan_launch_status();
}
extern "C" int spgan_gp_penalty_bwd(const float* g, const float* norms, const float* gout, int B, int64_t D,
                                    float gamma, float lambda, float* dg, spgan_stream_t s) {
    SPGAN_CHECK_ARG(g && norms && gout && dg && B >= 1 && D >= 1);
    gp_penalty_bwd_kernel<<<ew_grid((int64_t)B * D, 256), 256, 0, as_stream(s)>>>(g, norms, gout, B, D, gamma, lambda, dg);
    return spgan_launch_status();
}
extern "C" int spgan_mean(const float* x, int64_t n, float scale, int accumulate, float* out, spgan_stream_t s) {
    SPGAN_CHECK_ARG(x && out && n >= 1);
    mean_kernel<<<1, 1024, 0, as_stream(s)>>>(x, n, scale, accumulate, out);
    return spgan_launch_status();
}
