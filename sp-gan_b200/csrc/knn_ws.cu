// Warp-specialised variant of the fused kNN kernel (knn.cu::knn_group_fast_kernel): same arithmetic, same
// (dist, index) lists, bit-identical neighbour lists -- but the fp32 FMA phase and the selection phase of
// consecutive candidate tiles OVERLAP instead of alternating.
//
// Why (lead 2): per SM and per 64 x 128 tile the fast kernel spends 2.9 us in the FMA phase
// and 4.1 us in the selection phase; the selection is a chain of dependent shuffles / ballots (~160 warp
// instructions, ~1900 cycles per (query, tile)) that leaves the FMA pipe idle, and block barriers keep the two
// phases from overlapping.  Here a CTA of 768 threads owns an SM:
//   warps 0-7   FMA role (setmaxnreg 128): resident query tile, cp.async double-buffered candidate chunks, 4 x 8
//               register tiles, writes the 64 x 128 distance tile into one of TWO shared buffers;
//   warps 8-23  selection role (setmaxnreg 56): 4 queries per warp; copies its rows of the tile into registers,
//               releases the buffer at once, then runs the sorted-list insertion while the FMA warps are already
//               computing the next tile.
// Hand-over through named barriers (bar.arrive / bar.sync): full[buf] (FMA arrives, selection waits) and
// empty[buf] (selection arrives, FMA waits before overwriting).
//
// STATUS: opt-in (SPGAN_KNN_WS=1), written at the end of round 1 after the GPU budget was spent -- compiled, not yet
// run.  tests/test_gpu_knn_ws.py (skipped unless SPGAN_KNN_WS=1) checks it bit for bit against the default kernel.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int QT = 64, CT = 128, CK = 32, DPAD = 4, QMAXC = 128;
constexpr int FMA_THREADS = 256, SEL_THREADS = 512, WS_THREADS = FMA_THREADS + SEL_THREADS;
constexpr int BAR_FMA = 1, BAR_FULL = 2, BAR_EMPTY = 4;       // named barrier ids (0 = __syncthreads)

struct WsSmem {
    float q[QMAXC][QT];          // 32 KB, channel-major, resident
    float c[2][CK][CT];          // 32 KB, double-buffered candidate chunk
    float d[2][QT][CT + DPAD];   // 66 KB, double-buffered distance tile
    float xs_q[QT];
    float xs_c[2][CT];
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool lex_less(float d0, int j0, float d1, int j1) { return d0 < d1 || (d0 == d1 && j0 < j1); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1)
knn_ws_kernel(const float* __restrict__ x, const float* __restrict__ xs, int C, int N, int k, int32_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem& s = *reinterpret_cast<WsSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q_tiles = (N + QT - 1) / QT;
    const int b = blockIdx.x / q_tiles;
    const int i0 = (blockIdx.x % q_tiles) * QT;
    const float* xb = x + (int64_t)b * C * N;
    const float* xsb = xs + (int64_t)b * N;
    const int n_chunks = (C + CK - 1) / CK;
    const int n_tiles = (N + CT - 1) / CT;

    if (warp < FMA_THREADS / 32) {
        // ============================================================ FMA role
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        const int tx = tid & 15, ty = tid >> 4;
        const int total = n_tiles * n_chunks;
        for (int e = tid; e < C * (QT / 4); e += FMA_THREADS) {
            const int ch = e / (QT / 4), q4 = (e % (QT / 4)) * 4;
            cp_async16(&s.q[ch][q4], xb + (int64_t)ch * N + i0 + q4, i0 + q4 < N);
        }
        if (tid < QT) s.xs_q[tid] = (i0 + tid < N) ? xsb[i0 + tid] : 0.f;
        auto stage = [&](int it, int buf) {
            const int tile = it / n_chunks, chunk = it % n_chunks;
            const int j0 = tile * CT, c0 = chunk * CK;
            for (int e = tid; e < CK * (CT / 4); e += FMA_THREADS) {
                const int cc = e / (CT / 4), j4 = (e % (CT / 4)) * 4;
                const int ch = c0 + cc;
                cp_async16(&s.c[buf][cc][j4], xb + (int64_t)(ch < C ? ch : 0) * N + j0 + j4, ch < C && j0 + j4 < N);
            }
            // candidates beyond N get |x_j|^2 = +inf => dist = +inf: they never pass the selection filter
            if (chunk == 0 && tid < CT) s.xs_c[tile & 1][tid] = (j0 + tid < N) ? xsb[j0 + tid] : INFINITY;
        };
        stage(0, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");

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
            bar_sync(BAR_FMA, FMA_THREADS);          // chunk `it` visible to the FMA warps; buffer buf^1 is free
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

            // ---- tile complete: dist = (-2*dot + xs_i) + xs_j (modules.py:696,699) into distance buffer tile & 1
            const int db = tile & 1;
            if (tile >= 2) bar_sync(BAR_EMPTY + db, WS_THREADS);       // the selection warps have copied tile - 2 out
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float xq = s.xs_q[ty * 4 + a];
                float out[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int jj = (c < 4) ? tx * 4 + c : 64 + tx * 4 + (c - 4);
                    out[c] = fminf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, acc[a][c]), xq), s.xs_c[db][jj]), FLT_MAX);   // see knn.cu ord_key
                }
                *reinterpret_cast<float4*>(&s.d[db][ty * 4 + a][tx * 4]) = make_float4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<float4*>(&s.d[db][ty * 4 + a][64 + tx * 4]) = make_float4(out[4], out[5], out[6], out[7]);
            }
            __threadfence_block();
            bar_arrive(BAR_FULL + db, WS_THREADS);
        }
    } else {
        // ============================================================ selection role: 4 queries per warp
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int sw = warp - FMA_THREADS / 32;       // 0..15
        const int K1 = k + 1;
        float ld[4];
        int lj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { ld[u] = FLT_MAX; lj[u] = 0x7fffffff; }

        for (int tile = 0; tile < n_tiles; ++tile) {
            const int db = tile & 1;
            const int j0 = tile * CT;
            bar_sync(BAR_FULL + db, WS_THREADS);       // the FMA warps have written this tile
            float dd[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 dv = *reinterpret_cast<const float4*>(&s.d[db][sw * 4 + u][lane * 4]);
                dd[u][0] = dv.x; dd[u][1] = dv.y; dd[u][2] = dv.z; dd[u][3] = dv.w;
            }
            // rows are in registers: hand the buffer back (only if a later tile will reuse it)
            if (tile + 2 < n_tiles) bar_arrive(BAR_EMPTY + db, WS_THREADS);

#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jb = j0 + lane * 4;
                if (tile == 0) {
                    // bulk seed: warp bitonic sort of one candidate per lane by (dist, index), rank r -> lane r
                    float sd = dd[u][0];
                    int sj = jb;
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                        for (int st = k2 >> 1; st > 0; st >>= 1) {
                            const float od = __shfl_xor_sync(0xffffffffu, sd, st);
                            const int oj = __shfl_xor_sync(0xffffffffu, sj, st);
                            const bool keep_min = ((lane & st) == 0) == ((lane & k2) == 0);
                            const bool other_less = lex_less(od, oj, sd, sj);
                            if (keep_min == other_less) { sd = od; sj = oj; }
                        }
                    }
                    ld[u] = sd; lj[u] = sj;
                }
                float tau = __shfl_sync(0xffffffffu, ld[u], K1 - 1);
                int tauj = __shfl_sync(0xffffffffu, lj[u], K1 - 1);
                if (tile != 0) {
                    const float mn = fminf(fminf(dd[u][0], dd[u][1]), fminf(dd[u][2], dd[u][3]));
                    if (__ballot_sync(0xffffffffu, mn <= tau) == 0u) continue;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (tile == 0 && t == 0) continue;                  // already in the sorted seed
                    const int j = jb + t;
                    const float d = dd[u][t];
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
        }
        // ---- ranks 1..k
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + sw * 4 + u;
            if (lane >= 1 && lane < K1 && i < N) idx[((int64_t)b * N + i) * k + (lane - 1)] = lj[u] < N ? lj[u] : min(i, N - 1);
        }
    }
}

}  // namespace

// Opt-in switch read once: SPGAN_KNN_WS=1.
bool spgan_knn_ws_enabled() {
    static const bool on = [] { const char* e = getenv("SPGAN_KNN_WS"); return e && e[0] == '1'; }();
    return on;
}

// Same contract as the fast path of spgan_knn_group with ee == NULL: N % 4 == 0, N >= 128, C <= 128, k + 1 <= 32,
// x 16-byte aligned.  Returns SPGAN_E_UNSUPPORTED outside that envelope.
int spgan_knn_ws_launch(const float* x, const float* xs, int B, int C, int N, int k, int32_t* idx, cudaStream_t st) {
    if (!(N % 4 == 0 && N >= CT && C <= QMAXC && k + 1 <= 32 && (reinterpret_cast<uintptr_t>(x) & 15) == 0))
        return SPGAN_E_UNSUPPORTED;
    static_assert(sizeof(WsSmem) <= 200 * 1024, "one CTA per SM");
    cudaError_t e = cudaFuncSetAttribute(knn_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
    if (e != cudaSuccess) return (int)e;
    const int64_t grid = (int64_t)B * ((N + QT - 1) / QT);
    if (grid > 0x7fffffffLL) return SPGAN_E_UNSUPPORTED;
    knn_ws_kernel<<<(unsigned)grid, WS_THREADS, sizeof(WsSmem), st>>>(x, xs, C, N, k, idx);
    return spgan_launch_status();
}
