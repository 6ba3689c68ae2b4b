// Weight-gradient GEMM, second generation:  C[Mo,No] (+)= A^T B,  A = dY [K, Mo], B = X [K, No] row-major fp32,
// K = number of points / edges (1e5 .. 1e6), Mo / No = channel counts (autograd of every Conv1d(k=1) / Conv2d(1x1) /
// Linear of Generator.py:56-71,107-135 and Discriminator.py:55-94).
//
// Same arithmetic as gemm_ts.cu (FP16S split x = hi + 2^-11 lo, three kind::f16 MMAs per product, main and cross
// terms in separate fp32 TMEM accumulators) and the same machinery, turned by 90 degrees: the contraction index K is
// the ROW index of both operands in memory, so
//   * both fp32 tiles (64 rows of K x 128 channels) arrive by TMA (128-byte swizzle, out-of-bounds zero fill);
//   * the A converters read the staged tile COLUMN-wise -- thread = output row m, one LDS.32 per k, a warp touches 32
//     consecutive floats of one staged row: conflict-free -- split, and write hi / lo pairs into TENSOR MEMORY
//     (tcgen05.st); the MMA takes A from TMEM ("TS" form);
//   * the B converters read their tile column-wise too (thread = (channel n, 8 consecutive k)) and write 16-byte
//     chunks of the K-major, 128-byte-swizzled operand tile [W_hi ; W_lo] in shared memory (fence.proxy.async once
//     per k-block -- there are no outstanding global loads in these warps for the fence to wait on);
//   * one CTA = (128 output rows, up to 128 output columns, one chunk of K); it accumulates its chunk in TMEM in
//     chains of 4096 rows, adds every finished chain into ITS partial tile, and a second, tiny kernel adds the
//     chunk partials in a fixed order (deterministic -- the first-generation kernel, gemm_tc_tn.cu, flushed every
//     1024 rows with fp32 atomics).
// The previous kernel loaded both operands with scalar, transposing global loads from 16 producer warps (512 mbarrier
// arrivals per k-block) at the TF32 rate: 126 TFLOP/s on the critic's 1024 x 256 x 131072 gradient.
//
// TMEM map: [0,256) accumulators, column tile t at t * 128: (main 64 | cross 64); [256,384) A operand, two stages of
// (32 columns hi pairs | 32 columns lo pairs).
// Warps (18): 0-7 A converters (lane quarter = warp % 4, k half = warp / 4; they also drain the accumulators), 8-15 B
// converters, 16 MMA issuer, 17 TMA producer.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

using namespace tc;

constexpr int BM = 128, BNT = 64, BK = 64;            // output rows, output columns per column tile, k per block
constexpr int MAX_NT = 2;                             // column tiles per CTA
constexpr int WG_THREADS = 18 * 32;
constexpr int B_WARP0 = 8, MMA_WARP = 16, TMA_WARP = 17;
constexpr int STAGES = 2;
constexpr int SA_BYTES = 4 * BK * 128;                // staged fp32 A tile: 4 boxes of 64 rows x 32 floats = 32 KB
constexpr int SB_BYTES = 4 * BK * 128;                // staged fp32 B tile, same geometry
constexpr int STAGE_BYTES = SA_BYTES + SB_BYTES;      // 64 KB
constexpr int OPB_BYTES = MAX_NT * 2 * BNT * 128;     // B operand [W_hi ; W_lo] x 2 column tiles = 32 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGES * OPB_BYTES + 512 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_A0 = 256;

constexpr int BAR_S_FULL = 0;                         // [2] TMA landed (tx)
constexpr int BAR_S_EMPTY = 2;                        // [2] 16 converter warps have read the staged tiles
constexpr int BAR_TA_FULL = 4;                        // [2] A k-block in TMEM (8 warps)
constexpr int BAR_TA_EMPTY = 6;                       // [2] tcgen05.commit
constexpr int BAR_OB_FULL = 8;                        // [2] B operand tile written (8 warps)
constexpr int BAR_OB_EMPTY = 10;                      // [2] tcgen05.commit
constexpr int BAR_D_FULL = 12;                        // a chain of FLUSH_KB k-blocks is complete (commit)
constexpr int BAR_D_EMPTY = 13;                       // the 8 A-converter warps have drained the accumulators
constexpr int NUM_BARS = 14;
// The tensor core adds into the fp32 TMEM accumulator with truncation: the error grows with the length of the
// accumulation chain (~1.2e-9 relative per k).  Chains are therefore cut every FLUSH_KB k-blocks (4096 rows): the
// converter warps drain the accumulators and add them, in fp32 and by the same thread every time (deterministic),
// into the CTA's partial tile (L2-resident); the next chain starts from zero.
constexpr int FLUSH_KB = 64;

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
// two fp32 -> packed fp16 hi pair and 2^11-scaled fp16 lo pair (element 0 in the low half)
__device__ __forceinline__ void split2_f16s(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    float h0, h1;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}"
        : "=f"(h0), "=f"(h1) : "r"(hi));
    const float r0 = (v0 - h0) * 2048.f, r1 = (v1 - h1) * 2048.f;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// staged fp32 tile: 4 boxes (32 channels each) of 64 k-rows x 128 bytes, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ float staged(const unsigned char* tile, int k, int ch) {
    const unsigned char* p = tile + (ch >> 5) * (BK * 128) + k * 128 + ((((ch & 31) >> 2) ^ (k & 7)) << 4) + ((ch & 3) << 2);
    return *reinterpret_cast<const float*>(p);
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// Drain of a finished accumulation chain by one of the 8 A-converter warps: lane quarter q = warp % 4 (the TMEM lanes a
// warp may touch), `part` = 0..1 selects one half of the CTA's nt * 64 output columns.  Every output element is owned by
// one thread for the whole kernel: partial += chain is a plain, ordered read-modify-write (deterministic).
__device__ __forceinline__ void drain_chain(float* partial_tile_row, bool row_ok, int No, int n0, int nt, int q, int part,
                                            bool first) {
    const int chunks = nt * 2;                                        // 16-column chunks per warp: half of the CTA's nt * 64 columns
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int cc = 0; cc < chunks; ++cc) {
        const int c = (part * chunks + cc) * 16;                      // column inside the CTA's tile
        const uint32_t taddr = lane_addr + (uint32_t)((c >> 6) * 2 * BNT + (c & 63));
        uint32_t rv[16], rw[16];
        tmem_ld16_issue(taddr, rv);
        tmem_ld16_issue(taddr + BNT, rw);
        tmem_ld_wait();
        asm volatile("" : "+r"(rv[0]), "+r"(rv[1]), "+r"(rv[2]), "+r"(rv[3]), "+r"(rv[4]), "+r"(rv[5]), "+r"(rv[6]), "+r"(rv[7]),
                          "+r"(rv[8]), "+r"(rv[9]), "+r"(rv[10]), "+r"(rv[11]), "+r"(rv[12]), "+r"(rv[13]), "+r"(rv[14]),
                          "+r"(rv[15]) :: "memory");
        asm volatile("" : "+r"(rw[0]), "+r"(rw[1]), "+r"(rw[2]), "+r"(rw[3]), "+r"(rw[4]), "+r"(rw[5]), "+r"(rw[6]), "+r"(rw[7]),
                          "+r"(rw[8]), "+r"(rw[9]), "+r"(rw[10]), "+r"(rw[11]), "+r"(rw[12]), "+r"(rw[13]), "+r"(rw[14]),
                          "+r"(rw[15]) :: "memory");
        if (!row_ok) continue;
        const int col0 = n0 + c;
        if (col0 + 16 <= No && (No & 3) == 0) {
            float4 old[4];
            if (!first) {
#pragma unroll
                for (int j = 0; j < 4; ++j) old[j] = *reinterpret_cast<const float4*>(partial_tile_row + col0 + 4 * j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 o;
                o.x = fmaf(__uint_as_float(rw[4 * j]), 1.f / 2048.f, __uint_as_float(rv[4 * j]));
                o.y = fmaf(__uint_as_float(rw[4 * j + 1]), 1.f / 2048.f, __uint_as_float(rv[4 * j + 1]));
                o.z = fmaf(__uint_as_float(rw[4 * j + 2]), 1.f / 2048.f, __uint_as_float(rv[4 * j + 2]));
                o.w = fmaf(__uint_as_float(rw[4 * j + 3]), 1.f / 2048.f, __uint_as_float(rv[4 * j + 3]));
                if (!first) { o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w; }
                *reinterpret_cast<float4*>(partial_tile_row + col0 + 4 * j) = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col0 + j < No) {
                    const float o = fmaf(__uint_as_float(rw[j]), 1.f / 2048.f, __uint_as_float(rv[j]));
                    partial_tile_row[col0 + j] = first ? o : partial_tile_row[col0 + j] + o;
                }
        }
    }
}

struct WgParams {
    int Mo, No;
    int64_t K;
    int m_tiles, n_groups, k_chunks;
    int64_t rows_per_chunk;          // multiple of BK
    float* partial;                  // [k_chunks, Mo, No]
    int* status;
    const float* b_scale;            // prologue of the B operand, per channel n (both NULL: identity):
    const float* b_shift;            //   B'[k, n] = LeakyReLU_slope(B[k, n] * b_scale[n] + b_shift[n])
    float b_slope;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = tc::align_smem_1024(smem_raw);
    unsigned char* stg = smem;                                   // [STAGES][A 32 KB | B 32 KB]
    unsigned char* opb = stg + STAGES * STAGE_BYTES;             // [STAGES][2 column tiles x (hi 8 KB | lo 8 KB)]
    uint64_t* bars = reinterpret_cast<uint64_t*>(opb + STAGES * OPB_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int slot) { return bar0 + 8u * slot; };
    volatile int* vstatus = p.status;

    // unit -> (k chunk, m tile, n group): CTAs running side by side share a chunk's rows in L2
    const int unit = blockIdx.x;
    const int tiles = p.m_tiles * p.n_groups;
    const int chunk = unit / tiles, tile = unit % tiles;
    const int m0 = (tile / p.n_groups) * BM, n0 = (tile % p.n_groups) * (MAX_NT * BNT);
    const int nt = min(MAX_NT, (p.No - n0 + BNT - 1) / BNT);                     // column tiles of this CTA
    const int64_t kbeg = (int64_t)chunk * p.rows_per_chunk;
    const int64_t kend = kbeg + p.rows_per_chunk < p.K ? kbeg + p.rows_per_chunk : p.K;
    const int KB = (int)((kend - kbeg + BK - 1) / BK);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(BAR_S_FULL + s), 1);
            mbar_init(bar(BAR_S_EMPTY + s), 16);
            mbar_init(bar(BAR_TA_FULL + s), 8);
            mbar_init(bar(BAR_TA_EMPTY + s), 1);
            mbar_init(bar(BAR_OB_FULL + s), 8);
            mbar_init(bar(BAR_OB_EMPTY + s), 1);
        }
        mbar_init(bar(BAR_D_FULL), 1);
        mbar_init(bar(BAR_D_EMPTY), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    if (warp == TMA_WARP && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();                 // the CTA owns the SM and all 512 columns (see gemm_ts.cu)

    if (warp == TMA_WARP) {
        // ================================================================ TMA producer: both fp32 tiles of a k-block
        const uint32_t s0 = smem_u32(stg);
        // only the 32-channel boxes that hold real channels are fetched (the converters of the others read stale
        // shared memory into accumulator rows / columns that are never stored)
        const int boxes_a = min(4, (p.Mo - m0 + 31) / 32), boxes_b = min(4, (p.No - n0 + 31) / 32);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(bar(BAR_S_EMPTY + stage), phase ^ 1, vstatus);
            if (elect_one()) {
                const uint32_t dst = s0 + (uint32_t)(stage * STAGE_BYTES);
                const int krow = (int)(kbeg + (int64_t)kb * BK);
                mbar_expect_tx(bar(BAR_S_FULL + stage), (uint32_t)((boxes_a + boxes_b) * (BK * 128)));
                for (int b = 0; b < boxes_a; ++b)
                    tma_load_2d(dst + b * (BK * 128), &tmA, m0 + 32 * b, krow, bar(BAR_S_FULL + stage));
                for (int b = 0; b < boxes_b; ++b)
                    tma_load_2d(dst + SA_BYTES + b * (BK * 128), &tmB, n0 + 32 * b, krow, bar(BAR_S_FULL + stage));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer (warp-uniform, one elected lane)
        constexpr uint32_t idesc_main = make_idesc_f16(BM, 2 * BNT);
        constexpr uint32_t idesc_x = make_idesc_f16(BM, BNT);
        const uint32_t ob0 = smem_u32(opb);
        int stage = 0;
        uint32_t phase = 0, dphase = 0;
        for (int kb = 0; kb < KB; ++kb) {
            const int in_chain = kb % FLUSH_KB;
            if (in_chain == 0 && kb > 0) {                  // the previous chain must have been drained
                mbar_wait(bar(BAR_D_EMPTY), dphase, vstatus);
                dphase ^= 1;
            }
            mbar_wait(bar(BAR_TA_FULL + stage), phase, vstatus);
            mbar_wait(bar(BAR_OB_FULL + stage), phase, vstatus);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_base = TMEM_A0 + (uint32_t)(stage * 64);
                for (int t = 0; t < nt; ++t) {
                    const uint64_t bd = make_desc_sw128(ob0 + (uint32_t)(stage * OPB_BYTES + t * (2 * BNT * 128)));
                    const uint32_t d_main = (uint32_t)(t * 2 * BNT), d_x = d_main + BNT;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t a_hi = a_base + (uint32_t)(s * 8);
                        umma_ts(d_main, a_hi, bd + (uint64_t)(s * 2), idesc_main, (in_chain > 0 || s > 0) ? 1u : 0u);
                        umma_ts(d_x, a_hi + 32, bd + (uint64_t)(s * 2), idesc_x, 1u);
                    }
                }
                umma_commit(bar(BAR_TA_EMPTY + stage));
                umma_commit(bar(BAR_OB_EMPTY + stage));
                if (in_chain == FLUSH_KB - 1 || kb == KB - 1) umma_commit(bar(BAR_D_FULL));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp < B_WARP0) {
        // ================================================================ A converters: staged column m -> TMEM lane m
        const int q = warp & 3, h = warp >> 2;             // lane quarter, k half (32 k each)
        const int m = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int stage = 0;
        uint32_t phase = 0, dphase = 0;
        float* prow = p.partial + ((int64_t)chunk * p.Mo + (m0 + m)) * p.No;
        auto do_drain = [&](bool first) {
            mbar_wait(bar(BAR_D_FULL), dphase, vstatus);
            dphase ^= 1;
            tc_fence_after();
            drain_chain(prow, m0 + m < p.Mo, p.No, n0, nt, q, h, first);
            tc_fence_before();
            warp_arrive(bar(BAR_D_EMPTY), lane);
        };
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(bar(BAR_S_FULL + stage), phase, vstatus);
            const unsigned char* tileA = stg + stage * STAGE_BYTES;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = h * 32 + 2 * i;
                split2_f16s(staged(tileA, k, m), staged(tileA, k + 1, m), hi[i], lo[i]);
            }
            warp_arrive(bar(BAR_S_EMPTY + stage), lane);          // staged values are in registers
            mbar_wait(bar(BAR_TA_EMPTY + stage), phase ^ 1, vstatus);
            tc_fence_after();
            const uint32_t a_base = TMEM_A0 + (uint32_t)(stage * 64) + lane_addr;
            tmem_st16(a_base + (uint32_t)(h * 16), hi);
            tmem_st16(a_base + (uint32_t)(32 + h * 16), lo);
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(bar(BAR_TA_FULL + stage), lane);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            if (kb % FLUSH_KB == FLUSH_KB - 1 || kb == KB - 1) do_drain(kb < FLUSH_KB);
        }
    } else {
        // ================================================================ B converters: staged columns -> K-major operand tile
        const int bw = warp - B_WARP0;                      // 0..7
        int stage = 0;
        uint32_t phase = 0;
        // Prologue (the BatchNorm + LeakyReLU the forward applied to this operand inside its GEMM, BnActLinearTrain):
        // the activated input of a weight gradient is formed here, in the converter, instead of by a norm_apply pass
        // that writes it to HBM for this kernel to read back.  A thread's four tasks are the channels j * 32 + lane.
        // (Rows beyond K are zero-filled in BOTH operands; dY = 0 there, so what the prologue makes of them is moot.)
        const bool has_pro = p.b_scale != nullptr;
        const float slope = p.b_slope;
        float psc[4], psh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + j * 32 + lane;
            psc[j] = (has_pro && gn < p.No) ? __ldg(p.b_scale + gn) : 0.f;
            psh[j] = (has_pro && gn < p.No) ? __ldg(p.b_shift + gn) : 0.f;
        }
        for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(bar(BAR_S_FULL + stage), phase, vstatus);
            const unsigned char* tileB = stg + stage * STAGE_BYTES + SA_BYTES;
            // task = (channel n in 0..127, chunk c of 8 consecutive k): 1024 tasks, 4 per thread; a warp covers 32
            // consecutive channels of one chunk, so its 8 scalar reads per task walk 8 staged rows conflict-free
            uint4 vhi[4], vlo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int task = (bw * 4 + j) * 32 + lane;          // 0..1023
                const int n = task & 127, c = task >> 7;
                if ((n >> 6) >= nt) continue;                       // (warp-uniform) column tile not in this CTA
                uint32_t hh[4], ll[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float x0 = staged(tileB, c * 8 + 2 * e, n), x1 = staged(tileB, c * 8 + 2 * e + 1, n);
                    if (has_pro) {                                   // same expression as gemm_ts.cu's converter
                        x0 = fmaf(x0, psc[j], psh[j]); x0 = fmaxf(x0, x0 * slope);
                        x1 = fmaf(x1, psc[j], psh[j]); x1 = fmaxf(x1, x1 * slope);
                    }
                    split2_f16s(x0, x1, hh[e], ll[e]);
                }
                vhi[j] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                vlo[j] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
            warp_arrive(bar(BAR_S_EMPTY + stage), lane);
            mbar_wait(bar(BAR_OB_EMPTY + stage), phase ^ 1, vstatus);
            unsigned char* ob = opb + stage * OPB_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int task = (bw * 4 + j) * 32 + lane;
                const int n = task & 127, c = task >> 7;
                const int t = n >> 6, r = n & 63;                   // column tile, row inside it
                if (t >= nt) continue;
                unsigned char* base = ob + t * (2 * BNT * 128);
                const uint32_t off_hi = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(base + off_hi) = vhi[j];
                *reinterpret_cast<uint4*>(base + BNT * 128 + off_hi) = vlo[j];      // lo rows 64..127: same (row & 7)
            }
            fence_proxy_async();
            warp_arrive(bar(BAR_OB_FULL + stage), lane);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// C[m, n] (+)= sum over chunks of partial[c, m, n], chunks added in order (deterministic)
__global__ void wg_reduce_kernel(const float* __restrict__ partial, int k_chunks, int Mo, int No, float* __restrict__ C,
                                 int64_t ldc, int accumulate) {
    const int64_t total = (int64_t)Mo * No;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / No), n = (int)(i % No);
        float s = 0.f;
        for (int c = 0; c < k_chunks; ++c) s += __ldg(partial + (int64_t)c * total + i);
        float* cp = C + (int64_t)m * ldc + n;
        *cp = accumulate ? *cp + s : s;
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

struct WgPlan { int m_tiles, n_groups, k_chunks; int64_t rows_per_chunk; };
WgPlan make_plan(int64_t Mo, int No, int64_t K) {
    WgPlan pl;
    pl.m_tiles = (int)((Mo + BM - 1) / BM);
    pl.n_groups = (No + MAX_NT * BNT - 1) / (MAX_NT * BNT);
    const int tiles = pl.m_tiles * pl.n_groups;
    int chunks = kNumSMs / tiles;                       // one wave of CTAs
    if (chunks < 1) chunks = 1;
    const int64_t kblocks = (K + BK - 1) / BK;
    if (chunks > kblocks) chunks = (int)kblocks;
    pl.rows_per_chunk = ((kblocks + chunks - 1) / chunks) * BK;
    pl.k_chunks = (int)((K + pl.rows_per_chunk - 1) / pl.rows_per_chunk);
    return pl;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool spgan_gemm_wg_supported(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb) {
    return Mo >= 16 && Mo <= 65536 && No >= 16 && K >= 4096 && K < (1LL << 31) && (lda % 4) == 0 &&
           (ldb % 4) == 0 && aligned16(A) && aligned16(B) && encode_tiled_fn() != nullptr;
}

// bytes of the chunk-partial buffer (after a 256-byte status block)
size_t spgan_gemm_wg_workspace(int64_t Mo, int No, int64_t K) {
    const WgPlan pl = make_plan(Mo, No, K);
    return 256 + (size_t)pl.k_chunks * (size_t)Mo * (size_t)No * sizeof(float);
}

int spgan_gemm_wg(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                  int64_t ldc, int accumulate, void* workspace, cudaStream_t st, const float* b_scale, const float* b_shift,
                  float b_slope) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return SPGAN_E_UNSUPPORTED;
    const WgPlan pl = make_plan(Mo, No, K);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    CUtensorMap tmA, tmB;
    const cuuint32_t box[2] = {32, BK};
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Mo, (cuuint64_t)K};
        const cuuint64_t gstride[1] = {(cuuint64_t)lda * sizeof(float)};
        if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SPGAN_E_UNSUPPORTED;
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)No, (cuuint64_t)K};
        const cuuint64_t gstride[1] = {(cuuint64_t)ldb * sizeof(float)};
        if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(B), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SPGAN_E_UNSUPPORTED;
    }
    WgParams p;
    p.Mo = (int)Mo; p.No = No; p.K = K;
    p.m_tiles = pl.m_tiles; p.n_groups = pl.n_groups; p.k_chunks = pl.k_chunks; p.rows_per_chunk = pl.rows_per_chunk;
    p.partial = reinterpret_cast<float*>(ws + 256);
    p.status = reinterpret_cast<int*>(ws);
    p.b_scale = b_scale; p.b_shift = b_shift; p.b_slope = b_slope;
    cudaError_t e = cudaFuncSetAttribute(gemm_wg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int grid = pl.m_tiles * pl.n_groups * pl.k_chunks;
    gemm_wg_kernel<<<grid, WG_THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
    int rc = spgan_launch_status();
    if (rc != SPGAN_OK) return rc;
    wg_reduce_kernel<<<ew_grid(Mo * (int64_t)No, 256), 256, 0, st>>>(p.partial, pl.k_chunks, (int)Mo, No, C, ldc, accumulate);
    return spgan_launch_status();
}
