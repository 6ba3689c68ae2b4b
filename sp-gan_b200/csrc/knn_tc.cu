// kNN graph of a point-major feature matrix by FILTER AND REFINE: tensor cores find, for every
// query, a small candidate set that provably contains the reference's neighbours; the reference's exact fp32 recipe
// (oracle/knn_recipe.c: FMA chain over channels, two separate roundings, (dist, index) order) then ranks only those.
// Replaces Generation/modules.py:695-704 (bmm + sum + add + full sort + slice) for EdgeConv2's 64-channel graph; the
// result is bit-identical to the CUDA-core kernel (knn.cu), which stays the path for every other shape.
//
// Stage 1, knn_tc_filter_kernel -- the machinery of gemm_ts.cu with a selection epilogue.  Per 128-query tile of a
//   cloud: the fp32 query rows arrive by TMA, are split (x = hi + 2^-11 lo, fp16) and written to TENSOR MEMORY once;
//   the cloud's candidate tiles [X_hi ; X_lo] (64 points each, pre-split) stream through shared memory by TMA, TWICE:
//     pass 0: d_a(i,j) = (|x_i|^2 + |x_j|^2) - 2 <x_i, x_j> from the three-MMA product; every selector thread owns a
//             query (TMEM lane = query row) and keeps, in a branch-free sorted register list (min/max pairs: no warp
//             divergence, no shuffles), the k+1 smallest minima of groups of 8 candidates;
//             tau_a = the (k+1)-th smallest of them, an upper bound of the (k+1)-th smallest d_a of the query;
//     pass 1: the same d_a values again (same instructions on the same data: bit-identical); every candidate with
//             d_a <= tau_a + margin is appended to the query's candidate list.
//   margin = 1.2e-4 (|x_i|^2 + |tau_a|): with eps(i,j) <= 1.5e-5 (|x_i|^2 + |x_j|^2) bounding |d_a - d_exact| (fp32
//   FMA chain of the recipe: C u |x_i||x_j|, C <= 256; cascade norms; the split's 2^-22; fp32 accumulation) and
//   |x_j|^2 <= 2 |x_i|^2 + 2 d(i,j) for the points that matter, every j among the exact k+1 nearest satisfies
//   d_a(j) <= d_e(j) + eps <= max_{m in approx top} d_e(m) + eps <= tau_a + 2 eps_max <= tau_a + margin.
// Stage 2, knn_tc_refine_kernel -- one warp per query: exact recipe distance of each candidate (lane = candidate),
//   rank by (dist, index) with shuffles, ranks 1..k written.  Queries whose list overflowed (duplicate-heavy clouds),
//   came up short, or are non-finite fall back to an exact brute-force scan inside the same kernel.
//
// TMEM map / warp roles / barriers: as gemm_ts.cu (accumulator buffers [0,256), A operand [256,512); 0-7 converters,
// 8-15 selectors, 16 MMA issuer, 17 / 18 TMA producers).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <float.h>

namespace {

using namespace tc;

constexpr int BM = 128, BN = 64, BK = 64;
constexpr int SEL_WARP0 = 8, MMA_WARP = 16, TMA_A_WARP = 17, TMA_B_WARP = 18;
constexpr int KT_THREADS = 19 * 32;
constexpr int A_STAGES = 3, A_STAGE_BYTES = 2 * BM * 128;
constexpr int B_STAGES = 5, B_STAGE_BYTES = 2 * BN * 128;
constexpr int MAX_KB = 4;
constexpr int KL = 16;                                // list slots per selector thread (k + 1 <= 16)
constexpr int CAPH = 16;                              // candidates per (query, half)
constexpr int XCH_BYTES = 8 * 32 * KL * 4;            // list exchange between the two halves of a query
constexpr int MAX_N = 4096;                           // points per cloud (the candidate norms live in shared memory)
constexpr int XS_BYTES = MAX_N * 4;
constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + XCH_BYTES + XS_BYTES + 512 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t TMEM_COLS = 512, TMEM_A0 = 256;
constexpr float MARGIN_REL = 1.2e-4f;

constexpr int BAR_SA_FULL = 0, BAR_SA_EMPTY = BAR_SA_FULL + A_STAGES, BAR_TA_FULL = BAR_SA_EMPTY + A_STAGES,
              BAR_TA_EMPTY = BAR_TA_FULL + 4, BAR_SB_FULL = BAR_TA_EMPTY + 2, BAR_SB_EMPTY = BAR_SB_FULL + B_STAGES,
              BAR_D_FULL = BAR_SB_EMPTY + B_STAGES, BAR_D_EMPTY = BAR_D_FULL + 2, NUM_BARS = BAR_D_EMPTY + 2;
static_assert(NUM_BARS * 8 + 16 <= 512, "barrier block");

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

struct KtParams {
    int B, N, C, K1;                 // clouds, points per cloud, channels, k + 1
    const float* xs;                 // [B * N] squared norms (recipe arithmetic)
    int32_t* cand;                   // [B * N, 2, CAPH]
    int32_t* cnt;                    // [B * N, 2]   (> CAPH: the half overflowed)
    int* status;
};

template <int KLT>                                     // list slots actually maintained (k + 1 <= KLT <= KL)
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KtParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = tc::align_smem_1024(smem_raw);
    unsigned char* smA = smem;
    unsigned char* smB = smA + A_STAGES * A_STAGE_BYTES;
    float* xch = reinterpret_cast<float*>(smB + B_STAGES * B_STAGE_BYTES);
    float* xs_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(xch) + XCH_BYTES);          // [N] of the current cloud
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(xs_s) + XS_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int slot) { return bar0 + 8u * slot; };
    volatile int* vstatus = p.status;

    const int KB = (p.C + BK - 1) / BK;
    const int NREG = KB <= 2 ? 2 : 1;
    const int m_tiles = (int)(((int64_t)p.B * p.N) / BM);
    const int NT = p.N / BN;                               // candidate tiles per cloud
    const int tiles_per_cloud = p.N / BM;

    if (tid == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(bar(BAR_SA_FULL + s), 1); mbar_init(bar(BAR_SA_EMPTY + s), 8); }
        for (int s = 0; s < 4; ++s) mbar_init(bar(BAR_TA_FULL + s), 8);
        for (int s = 0; s < 2; ++s) mbar_init(bar(BAR_TA_EMPTY + s), 1);
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(bar(BAR_SB_FULL + s), 1); mbar_init(bar(BAR_SB_EMPTY + s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar(BAR_D_FULL + s), 1); mbar_init(bar(BAR_D_EMPTY + s), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    if (warp == TMA_A_WARP && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    if (warp == TMA_B_WARP && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();                          // see gemm_ts.cu

    if (warp == TMA_A_WARP) {
        // ================================================================ TMA producer: fp32 query rows
        const uint32_t sA0 = smem_u32(smA);
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(bar(BAR_SA_EMPTY + stage), phase ^ 1, vstatus);
                if (elect_one()) {
                    const uint32_t dst = sA0 + (uint32_t)(stage * A_STAGE_BYTES);
                    mbar_expect_tx(bar(BAR_SA_FULL + stage), A_STAGE_BYTES);
                    tma_load_2d(dst, &tmA, kb * BK, mt * BM, bar(BAR_SA_FULL + stage));
                    tma_load_2d(dst + BM * 128, &tmA, kb * BK + 32, mt * BM, bar(BAR_SA_FULL + stage));
                }
                __syncwarp();
                if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == TMA_B_WARP) {
        // ================================================================ TMA producer: the cloud's candidate tiles, twice
        const uint32_t sB0 = smem_u32(smB);
        int stage = 0;
        uint32_t phase = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            const int tile0 = (mt / tiles_per_cloud) * NT;
            for (int it = 0; it < 2 * NT; ++it) {
                const int nt = tile0 + (it >= NT ? it - NT : it);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar(BAR_SB_EMPTY + stage), phase ^ 1, vstatus);
                    if (elect_one()) {
                        mbar_expect_tx(bar(BAR_SB_FULL + stage), B_STAGE_BYTES);
                        tma_load_2d(sB0 + (uint32_t)(stage * B_STAGE_BYTES), &tmB, kb * BK, nt * 2 * BN, bar(BAR_SB_FULL + stage));
                    }
                    __syncwarp();
                    if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer (warp-uniform, one elected lane)
        constexpr uint32_t idesc_main = make_idesc_f16(BM, 2 * BN);
        constexpr uint32_t idesc_x = make_idesc_f16(BM, BN);
        const uint32_t sB0 = smem_u32(smB);
        int sb = 0, buf = 0;
        uint32_t sb_phase = 0, d_phase = 0;
        uint32_t ta_phase[2] = {0, 0};
        int iter = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++iter) {
            const int g = NREG == 2 ? (iter & 1) : 0;
            const uint32_t a_base = TMEM_A0 + (uint32_t)(g * 128);
            for (int it = 0; it < 2 * NT; ++it) {
                mbar_wait(bar(BAR_D_EMPTY + buf), d_phase ^ 1, vstatus);
                tc_fence_after();
                const uint32_t d_main = (uint32_t)(buf * 2 * BN);
                const uint32_t d_x = d_main + BN;
                for (int kb = 0; kb < KB; ++kb) {
                    if (it == 0) mbar_wait(bar(BAR_TA_FULL + g * 2 * (NREG - 1) + kb), ta_phase[g], vstatus);
                    mbar_wait(bar(BAR_SB_FULL + sb), sb_phase, vstatus);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bd = make_desc_sw128(sB0 + (uint32_t)(sb * B_STAGE_BYTES));
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const uint32_t a_hi = a_base + (uint32_t)(kb * 64 + s * 8);
                            umma_ts(d_main, a_hi, bd + (uint64_t)(s * 2), idesc_main, (kb > 0 || s > 0) ? 1u : 0u);
                            umma_ts(d_x, a_hi + 32, bd + (uint64_t)(s * 2), idesc_x, 1u);
                        }
                        umma_commit(bar(BAR_SB_EMPTY + sb));
                    }
                    __syncwarp();
                    if (++sb == B_STAGES) { sb = 0; sb_phase ^= 1; }
                }
                if (elect_one()) {
                    umma_commit(bar(BAR_D_FULL + buf));
                    if (it == 2 * NT - 1) umma_commit(bar(BAR_TA_EMPTY + g));
                }
                __syncwarp();
                buf ^= 1;
                if (buf == 0) d_phase ^= 1;
            }
            ta_phase[g] ^= 1;
        }
    } else if (warp < SEL_WARP0) {
        // ================================================================ converters: staged fp32 rows -> hi / lo -> TMEM
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int stage = 0;
        uint32_t sa_phase = 0;
        uint32_t te_phase[2] = {0, 0};
        int iter = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++iter) {
            const int g = NREG == 2 ? (iter & 1) : 0;
            if (iter >= NREG) {
                mbar_wait(bar(BAR_TA_EMPTY + g), te_phase[g], vstatus);
                te_phase[g] ^= 1;
                tc_fence_after();
            }
            const uint32_t a_base = TMEM_A0 + (uint32_t)(g * 128) + lane_addr;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(bar(BAR_SA_FULL + stage), sa_phase, vstatus);
                const unsigned char* box = smA + stage * A_STAGE_BYTES + h * (BM * 128) + row * 128;
                float v[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 t = *reinterpret_cast<const float4*>(box + ((c ^ (row & 7)) << 4));
                    v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
                    float h0, h1;
                    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}"
                        : "=f"(h0), "=f"(h1) : "r"(hi[i]));
                    const float r0 = (v[2 * i] - h0) * 2048.f, r1 = (v[2 * i + 1] - h1) * 2048.f;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo[i]) : "f"(r1), "f"(r0));
                }
                warp_arrive(bar(BAR_SA_EMPTY + stage), lane);
                tmem_st16(a_base + (uint32_t)(kb * 64 + h * 16), hi);
                tmem_st16(a_base + (uint32_t)(kb * 64 + 32 + h * 16), lo);
                tmem_st_wait();
                tc_fence_before();
                warp_arrive(bar(BAR_TA_FULL + g * 2 * (NREG - 1) + kb), lane);
                if (++stage == A_STAGES) { stage = 0; sa_phase ^= 1; }
            }
        }
    } else {
        // ================================================================ selectors: thread = (query, half of each tile)
        const int ew = warp - SEL_WARP0;
        const int q = ew & 3, h = ew >> 2;
        float* my_x = xch + (ew * 32 + lane) * KL;
        const float* peer_x = xch + ((ew ^ 4) * 32 + lane) * KL;
        const int K1 = p.K1;
        int buf = 0, cur_cloud = -1;
        uint32_t d_phase = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            const int64_t qrow = (int64_t)mt * BM + q * 32 + lane;          // global point index of this thread's query
            const int cloud = mt / tiles_per_cloud;
            const float xs_i = __ldg(p.xs + qrow);
            if (cloud != cur_cloud) {
                // the cloud's candidate norms -> shared memory (all 8 selector warps; every one of them has finished
                // the previous tile when it gets here, the named barrier makes that mutual)
                asm volatile("bar.sync 5, 256;" ::: "memory");
                const float4* src = reinterpret_cast<const float4*>(p.xs + (int64_t)cloud * p.N);
                for (int i = (ew * 32 + lane); i < p.N / 4; i += 256) reinterpret_cast<float4*>(xs_s)[i] = __ldg(src + i);
                asm volatile("bar.sync 5, 256;" ::: "memory");
                cur_cloud = cloud;
            }
            float L[KLT];
#pragma unroll
            for (int r = 0; r < KLT; ++r) L[r] = FLT_MAX;
            float thr = 0.f;
            int count = 0;
            int32_t* my_cand = p.cand + (qrow * 2 + h) * CAPH;
            for (int it = 0; it < 2 * NT; ++it) {
                const int nt = it >= NT ? it - NT : it;
                const int j0 = nt * BN + h * 32;                             // first candidate (inside the cloud) of this thread
                mbar_wait(bar(BAR_D_FULL + buf), d_phase, vstatus);
                tc_fence_after();
                const uint32_t taddr = ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN + h * 32);
                float d[32];
                {
                    uint32_t rv[32], rw[32];
                    tmem_ld32_issue(taddr, rv);
                    tmem_ld32_issue(taddr + BN, rw);
                    tmem_ld_wait();
                    tmem_pin32(rv);
                    tmem_pin32(rw);
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = fmaf(__uint_as_float(rw[j]), 1.f / 2048.f, __uint_as_float(rv[j]));
                }
                tc_fence_before();
                warp_arrive(bar(BAR_D_EMPTY + buf), lane);
                buf ^= 1;
                if (buf == 0) d_phase ^= 1;
                // d_a = (|x_i|^2 + |x_j|^2) - 2 dot: the same instructions in both passes => the same values
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 xj = reinterpret_cast<const float4*>(xs_s + j0)[c];
                    d[4 * c] = fmaf(-2.f, d[4 * c], xs_i + xj.x);
                    d[4 * c + 1] = fmaf(-2.f, d[4 * c + 1], xs_i + xj.y);
                    d[4 * c + 2] = fmaf(-2.f, d[4 * c + 2], xs_i + xj.z);
                    d[4 * c + 3] = fmaf(-2.f, d[4 * c + 3], xs_i + xj.w);
                }
                if (it < NT) {
                    // ---- pass 0: an UPPER BOUND of the (k+1)-th smallest d_a is all the threshold needs.  The list holds
                    // the smallest MINIMA OF GROUPS of 8 candidates: K1 group minima <= t means K1 distinct candidates
                    // <= t, so the K1-th smallest group minimum bounds the K1-th smallest candidate (and equals it unless
                    // two of the K1 nearest share a group).  1 + KLT / 4 min/max per candidate instead of 2 KLT, branch
                    // free: no warp divergence, no shuffles.  (NaN never enters: fminf / fmaxf drop it.)
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        float v = fminf(fminf(fminf(d[8 * g8], d[8 * g8 + 1]), fminf(d[8 * g8 + 2], d[8 * g8 + 3])),
                                        fminf(fminf(d[8 * g8 + 4], d[8 * g8 + 5]), fminf(d[8 * g8 + 6], d[8 * g8 + 7])));
                        v = fminf(v, FLT_MAX);
#pragma unroll
                        for (int r = 0; r < KLT; ++r) {
                            const float lo = fminf(L[r], v);
                            v = fmaxf(L[r], v);
                            L[r] = lo;
                        }
                    }
                    if (it == NT - 1) {
                        // ---- merge the two halves of the query, form the threshold
#pragma unroll
                        for (int r = 0; r < KLT; ++r) my_x[r] = L[r];
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
#pragma unroll
                        for (int r2 = 0; r2 < KLT; ++r2) {
                            float v = peer_x[r2];
#pragma unroll
                            for (int r = 0; r < KLT; ++r) {
                                const float lo = fminf(L[r], v);
                                v = fmaxf(L[r], v);
                                L[r] = lo;
                            }
                        }
                        float tau = L[0];
#pragma unroll
                        for (int r = 1; r < KLT; ++r) tau = (r == K1 - 1) ? L[r] : tau;
                        thr = tau + MARGIN_REL * (xs_i + fabsf(tau));
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // lists read: the slots may be reused
                    }
                } else {
                    // ---- pass 1: collect every candidate under the threshold (bit mask first: one short loop per hit)
                    uint32_t hits = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) hits |= (d[j] <= thr) ? (1u << j) : 0u;
                    while (hits) {
                        const int j = __ffs(hits) - 1;
                        hits &= hits - 1;
                        if (count < CAPH) my_cand[count] = j0 + j;
                        ++count;
                    }
                }
            }
            p.cnt[qrow * 2 + h] = count;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// point-major rows [R, C] -> fp16 hi / 2^11-scaled lo, per 64-point tile [hi rows ; lo rows] x Kp (one TMA box per tile)
__global__ void knn_presplit_kernel(const float* __restrict__ rows, int64_t R, int C, int Kp, uint16_t* __restrict__ out) {
    const int64_t total = R * Kp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Kp;
        const int k = (int)(i % Kp);
        const float v = k < C ? __ldg(rows + r * C + k) : 0.f;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
        const int64_t t = r / BN;
        const int rr = (int)(r % BN);
        out[(t * 2 * BN + rr) * Kp + k] = __half_as_ushort(hi);
        out[(t * 2 * BN + BN + rr) * Kp + k] = __half_as_ushort(lo);
    }
}

__device__ __forceinline__ bool key_less(float d0, int j0, float d1, int j1) { return d0 < d1 || (d0 == d1 && j0 < j1); }

// the reference's exact fp32 distance (modules.py:696-699 as torch CPU rounds it: oracle/knn_recipe.c)
__device__ __forceinline__ float recipe_dist(const float* __restrict__ xi, const float* __restrict__ xj, int C, float xs_i,
                                             float xs_j) {
    float acc = 0.f;
    int c = 0;
    // 32 channels per step: all 16 loads of the step are issued before the (sequential, channel-order) FMA chain
    // consumes them -- the candidate rows are scattered over L2, a dependent load per FMA group costs a round trip
    for (; c + 32 <= C; c += 32) {
        float4 a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a[u] = __ldg(reinterpret_cast<const float4*>(xi + c) + u);
            b[u] = __ldg(reinterpret_cast<const float4*>(xj + c) + u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc = __fmaf_rn(a[u].x, b[u].x, acc);
            acc = __fmaf_rn(a[u].y, b[u].y, acc);
            acc = __fmaf_rn(a[u].z, b[u].z, acc);
            acc = __fmaf_rn(a[u].w, b[u].w, acc);
        }
    }
    for (; c < C; c += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xi + c)), b = __ldg(reinterpret_cast<const float4*>(xj + c));
        acc = __fmaf_rn(a.x, b.x, acc);
        acc = __fmaf_rn(a.y, b.y, acc);
        acc = __fmaf_rn(a.z, b.z, acc);
        acc = __fmaf_rn(a.w, b.w, acc);
    }
    return fminf(__fadd_rn(__fadd_rn(__fmul_rn(-2.0f, acc), xs_i), xs_j), FLT_MAX);      // ord_key of knn.cu
}

// one warp per query
__global__ void __launch_bounds__(256, 2)
knn_tc_refine_kernel(const float* __restrict__ rows, const float* __restrict__ xs, const int32_t* __restrict__ cand,
                     const int32_t* __restrict__ cnt, int B, int N, int C, int k, int32_t* __restrict__ idx,
                     int* __restrict__ fallbacks) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)B * N;
    const int K1 = k + 1;
    for (int64_t qrow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); qrow < total;
         qrow += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int cloud = (int)(qrow / N);
        const int64_t base = (int64_t)cloud * N;
        const float* xi = rows + qrow * C;
        const float xs_i = __ldg(xs + qrow);
        const int c0 = __ldg(cnt + qrow * 2), c1 = __ldg(cnt + qrow * 2 + 1);
        int32_t* out = idx + qrow * k;
        if (c0 <= CAPH && c1 <= CAPH && c0 + c1 >= K1) {
            // ---- the common case: rank the candidates exactly
            const int n = c0 + c1;                                     // <= 32
            int j = 0x7fffffff;
            float d = FLT_MAX;
            if (lane < n) {
                j = lane < c0 ? __ldg(cand + qrow * 2 * CAPH + lane) : __ldg(cand + (qrow * 2 + 1) * CAPH + (lane - c0));
                d = recipe_dist(xi, rows + (base + j) * C, C, xs_i, __ldg(xs + base + j));
            }
            int rank = 0;
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const float dm = __shfl_sync(0xffffffffu, d, m);
                const int jm = __shfl_sync(0xffffffffu, j, m);
                rank += key_less(dm, jm, d, j) ? 1 : 0;
            }
            if (lane < n && rank >= 1 && rank <= k) out[rank - 1] = j;
            continue;
        }
        // ---- fallback: exact scan of the whole cloud, k + 1 rounds of "smallest key above the last one"
        if (lane == 0) atomicAdd(fallbacks, 1);
        float last_d = -FLT_MAX;
        int last_j = -1;
        bool first = true;
        for (int r = 0; r < K1; ++r) {
            float bd = FLT_MAX;
            int bj = 0x7fffffff;
            for (int j = lane; j < N; j += 32) {
                const float d = recipe_dist(xi, rows + (base + j) * C, C, xs_i, __ldg(xs + base + j));
                const bool above = first || key_less(last_d, last_j, d, j);
                if (above && key_less(d, j, bd, bj)) { bd = d; bj = j; }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (key_less(od, oj, bd, bj)) { bd = od; bj = oj; }
            }
            last_d = bd; last_j = bj; first = false;
            if (r >= 1 && lane == 0) out[r - 1] = bj < N ? bj : min((int)(qrow - base), N - 1);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" size_t spgan_knn_rows_workspace(int B, int C, int N, int k) {
    if (B < 1 || N < BM || N % BM != 0 || C < 4 || C > MAX_KB * BK || C % 4 != 0 || k < 1 || k + 1 > KL || k + 1 > N ||
        N > MAX_N || (int64_t)B * N >= (1LL << 31) || encode_tiled_fn() == nullptr)
        return 0;
    const size_t R = (size_t)B * N, Kp = align_up((size_t)C, BK);
    return 256 + align_up(R * 2 * Kp * sizeof(uint16_t), 256) + align_up(R * 2 * CAPH * sizeof(int32_t), 256) +
           align_up(R * 2 * sizeof(int32_t), 256);
}

extern "C" int spgan_knn_rows(const float* rows, const float* xs, int B, int C, int N, int k, int32_t* idx, void* workspace,
                              size_t workspace_bytes, spgan_stream_t stream) {
    SPGAN_CHECK_ARG(rows && xs && idx && workspace);
    const size_t need = spgan_knn_rows_workspace(B, C, N, k);
    if (need == 0 || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0 ||
        (reinterpret_cast<uintptr_t>(rows) & 15) != 0 || (reinterpret_cast<uintptr_t>(xs) & 15) != 0)
        return SPGAN_E_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    EncodeTiledFn enc = encode_tiled_fn();
    const int64_t R = (int64_t)B * N;
    const int Kp = (int)align_up((size_t)C, BK);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int* status = reinterpret_cast<int*>(ws);                      // [0] pipeline status, [1] fallback counter
    uint16_t* split = reinterpret_cast<uint16_t*>(ws + 256);
    int32_t* cand = reinterpret_cast<int32_t*>(ws + 256 + align_up((size_t)R * 2 * Kp * sizeof(uint16_t), 256));
    int32_t* cnt = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(cand) + align_up((size_t)R * 2 * CAPH * sizeof(int32_t), 256));
    cudaError_t e = cudaMemsetAsync(status, 0, 256, st);
    if (e != cudaSuccess) return (int)e;
    knn_presplit_kernel<<<ew_grid(R * Kp, 256), 256, 0, st>>>(rows, R, C, Kp, split);
    int rc = spgan_launch_status();
    if (rc != SPGAN_OK) return rc;
    CUtensorMap tmA, tmB;
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
        const cuuint64_t gstride[1] = {(cuuint64_t)C * sizeof(float)};
        const cuuint32_t box[2] = {32, BM};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(rows), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SPGAN_E_UNSUPPORTED;
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)(R / BN) * 2 * BN};
        const cuuint64_t gstride[1] = {(cuuint64_t)Kp * sizeof(uint16_t)};
        const cuuint32_t box[2] = {BK, 2 * BN};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, split, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
            CUDA_SUCCESS)
            return SPGAN_E_UNSUPPORTED;
    }
    KtParams p;
    p.B = B; p.N = N; p.C = C; p.K1 = k + 1; p.xs = xs; p.cand = cand; p.cnt = cnt; p.status = status;
    const int m_tiles = (int)(R / BM);
    const int grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
    if (k + 1 <= 12) {
        e = cudaFuncSetAttribute(knn_tc_filter_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        knn_tc_filter_kernel<12><<<grid, KT_THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
    } else {
        e = cudaFuncSetAttribute(knn_tc_filter_kernel<KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        knn_tc_filter_kernel<KL><<<grid, KT_THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
    }
    rc = spgan_launch_status();
    if (rc != SPGAN_OK) return rc;
    knn_tc_refine_kernel<<<kNumSMs * 16, 256, 0, st>>>(rows, xs, cand, cnt, B, N, C, k, idx, status + 1);
    return spgan_launch_status();
}
