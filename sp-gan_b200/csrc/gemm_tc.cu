// tcgen05 GEMM (engine 1 of spgan_gemm): C[M,N] = A[M,K] * B^T (+bias) (+C), fp32 in / fp32 out.
//
// Precision: every fp32 operand is split into two narrow terms x = hi + lo and the product is formed
// as Ahi*Bhi + Ahi*Blo + Alo*Bhi on the 5th-gen tensor cores with fp32 accumulation in TMEM:
//   TF32 mode (engine 1, default): hi = rna_tf32(x), lo = rna_tf32(x - hi) (integer rounding, 2 full-rate
//       ops each -- cvt.rna.tf32 is a 16-lane/clk instruction): 22 significant bits,
//       ~2^-21 relative per product -- indistinguishable from an fp32 FMA chain at K <= 1280;
//   BF16 mode (engine 2, "fast"): hi = trunc_bf16(x), lo = rn_bf16(x - hi): ~2^-16 per product,
//       twice the MMA rate; NOT used by default because train-mode BatchNorm and the gradient
//       penalty amplify it to the 1e-3 parity bar (DESIGN.md "GEMM precision").
//   FP16S mode (engine 3): hi = rn_f16(x), lo = rn_f16((x - hi) * 2^11): fp16 carries the same 11 significant
//       bits as tf32, so the split is as fine as TF32 mode (22 bits) at the kind::f16 MMA rate (2x tf32).  The
//       residual is scaled by 2^11 so that it stays a NORMAL fp16 number whatever |x| is; the cross terms
//       therefore accumulate 2^11 too large in their own TMEM accumulator and the epilogue forms
//       main + 2^-11 * cross.  Needs |x| < 65504 (fp16 range).  Validated on B200 in round 2
//       (tests/test_gpu_gemm_tc.py: same error class as TF32x3 against fp64).
//
// Structure (one persistent CTA per SM, 20 warps):
//   warps 0-7   epilogue : tcgen05.ld accumulator rows from TMEM -> shared transpose -> (+bias, +C)
//                          -> 128-byte coalesced global stores (two warps per TMEM lane quarter)
//   warp  8     MMA      : one lane issues tcgen05.mma (M=128, N=BN, K=8 tf32 / 16 bf16) and commits
//   warps 9-16  A producer: global fp32 A tile -> split -> 128B-swizzled K-major shared tiles, with
//                          the next k-blocks' loads in flight during conversion.  (No TMA: the A
//                          operand needs the fp32 -> hi/lo conversion on the way in.)
//   warps 17-19 B loader : pre-split weight tiles -> shared with cp.async (from L2), two stages in flight
// Pipelines: full/empty mbarriers per shared stage, tmem_full/tmem_empty per accumulator buffer
// (two buffers, so the epilogue of tile i overlaps the MMAs of tile i+1).
// Every mbarrier wait is bounded; on timeout the kernel raises a status word instead of hanging.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
// one k-block = one 128-byte swizzle row per matrix row: 64 bf16 or 32 tf32 elements
constexpr int NUM_EPI_WARPS = 8;            // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int MMA_WARP = 8;
constexpr int A_WARP0 = 9;                  // 8 warps convert the fp32 A operand
constexpr int NUM_A_THREADS = 8 * 32;
constexpr int B_WARP0 = 17;                 // 3 warps stream the pre-split B operand (cp.async, 2 stages in flight)
constexpr int NUM_B_THREADS = 3 * 32;
constexpr int TC_THREADS = 20 * 32;         // 640: the register file allows 96 registers per thread
constexpr int EPI_LD = 32;                  // epilogue staging tile: 32 x 32 floats per warp, XOR-swizzled float4 columns
#ifndef SPGAN_TC_PF_DEFAULT
#define SPGAN_TC_PF_DEFAULT 3
#endif
#ifndef SPGAN_TC_CONSUMER_FENCE_DEFAULT
#define SPGAN_TC_CONSUMER_FENCE_DEFAULT 0
#endif

enum : int { MODE_TF32 = 0, MODE_BF16 = 1, MODE_F16S = 2 };

template <int BN, int MODE>
struct Cfg {
    static constexpr bool TF32 = MODE == MODE_TF32;
    static constexpr bool TWOACC = MODE != MODE_BF16;    // separate accumulator for the cross terms
    static constexpr int BK = TF32 ? 32 : 64;            // elements per k-block
    static constexpr int CHUNK = TF32 ? 4 : 8;           // elements per 16-byte chunk
    static constexpr int A_BYTES = BM * 128;             // one half (hi or lo)
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
    // TF32 mode keeps the small cross terms (hi*lo + lo*hi) in a second accumulator: the tensor core adds
    // into TMEM with truncation, so the error grows with the number of accumulations into the LARGE sum;
    // this cuts that count by 3 (measured: rms error ~7e-9*K -> ~2.4e-9*K).
    static constexpr int NACC = TWOACC ? 2 : 1;
    static constexpr int TMEM_COLS = 2 * NACC * BN;      // two tile buffers; power of two <= 512
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * 32 * EPI_LD * 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                      EPI_BYTES /*epilogue staging*/;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

using namespace tc;

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32 [4,6), A format [7,10), B format [10,13) (1 = bf16, 2 = tf32), both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int mode) {
    const uint32_t fmt = mode == MODE_TF32 ? 2u : (mode == MODE_F16S ? 0u : 1u);    // kind::f16 formats: 0 = f16, 1 = bf16
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t swz(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// 8 fp32 -> 8 bf16 "hi" (truncation) + 8 bf16 "lo" (rn of the remainder), element 0 in the low half
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t u0 = __float_as_uint(v[2 * i]), u1 = __float_as_uint(v[2 * i + 1]);
        h[i] = __byte_perm(u0, u1, 0x7632);
        const float r0 = v[2 * i] - __uint_as_float(u0 & 0xffff0000u);
        const float r1 = v[2 * i + 1] - __uint_as_float(u1 & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(r1), "f"(r0));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 8 fp32 -> 8 fp16 "hi" (round to nearest) + 8 fp16 "lo" = rn((x - hi) * 2^11), element 0 in the low half
__device__ __forceinline__ void split8_f16s(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
        float h0, h1;
        asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}"
            : "=f"(h0), "=f"(h1) : "r"(h[i]));
        const float r0 = (v[2 * i] - h0) * 2048.f, r1 = (v[2 * i + 1] - h1) * 2048.f;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(r1), "f"(r0));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// CF ("consumer fence"): where the generic->async proxy fence for the operand tiles is executed.  false: by every
// producer thread before its mbarrier arrive (the textbook placement).  true: once per k-block by the MMA-issuing
// thread after its acquire of the full barrier.  fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and
// the MEMBAR makes a producer wait for ITS OWN outstanding global prefetch loads: with the fence on the producer
// side every k-block costs a full memory latency no matter how deep the prefetch ring is.
template <int BN, int MODE, int PF, bool CF>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(int64_t M, int N, int K, const float* __restrict__ A, int64_t lda,
               const unsigned char* __restrict__ Bhi, const unsigned char* __restrict__ Blo, int Kp, int n_tiles,
               float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int accumulate, int* status,
               bool vecA, bool vecC) {
    using cfg = Cfg<BN, MODE>;
    constexpr bool TF32 = cfg::TF32, TWOACC = cfg::TWOACC;
    constexpr int BK = cfg::BK, CHUNK = cfg::CHUNK, ESZ = TF32 ? 4 : 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = tc::align_smem_1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE_BYTES);
    // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem base slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * cfg::STAGES + 4);
    float* epi_smem = reinterpret_cast<float*>(smem + cfg::STAGES * cfg::STAGE_BYTES + 256);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (see tc::elect_one)
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (cfg::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * cfg::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * cfg::STAGES + 2 + a); };

    if (tid == 0) {
        for (int s = 0; s < cfg::STAGES; ++s) {
            mbar_init(full_bar(s), NUM_A_THREADS + NUM_B_THREADS);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), NUM_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int m_tiles = (int)((M + BM - 1) / BM);
    const int KB = Kp / BK;
    volatile int* vstatus = status;
    // persistent tile walk t = blockIdx.x, += gridDim.x over (m_tile, n_tile) pairs, n fastest: kept as two
    // small integers updated incrementally -- no 64-bit divisions inside the role loops
    const int g_div = (int)gridDim.x / n_tiles, g_mod = (int)gridDim.x % n_tiles;
    const int mt0 = (int)blockIdx.x / n_tiles, nt0 = (int)blockIdx.x % n_tiles;
    auto tile_next = [&](int& mt, int& nt) { nt += g_mod; mt += g_div; if (nt >= n_tiles) { nt -= n_tiles; ++mt; } };

    if (warp >= A_WARP0 && warp < B_WARP0) {
        // ================================================================ A producers (fp32 -> bf16 hi/lo)
        // Each thread owns 4 (row, 8-float chunk) tasks of a k-block; the loads of the NEXT k-block are
        // issued before the current one is converted, so HBM latency overlaps conversion and the wait
        // for a free stage.
        constexpr int TASKS = BM * 8 / NUM_A_THREADS;          // 4 (row, 16-byte chunk) tasks per thread
        constexpr int V = TF32 ? 1 : 2;                        // float4 loads per task
        const int ptid = tid - A_WARP0 * 32;
        // PF register sets form a ring: PF-1 k-blocks of loads are in flight while one is converted.  (The first
        // version kept two named sets and copied nxt -> cur at the end of every iteration: that copy waits for
        // the loads it reads, so the NEXT loads could only be issued after the previous ones had landed -- one
        // k-block in flight per SM, ~1 TB/s of A traffic on K <= 128 shapes.)
        float4 buf[PF][V * TASKS];
        uint32_t soff[TASKS];
        int colo[TASKS];
        const float* rowp[TASKS];                              // row pointers of the tile being LOADED
#pragma unroll
        for (int j = 0; j < TASKS; ++j) {
            const int task = ptid + j * NUM_A_THREADS;
            soff[j] = swz(task >> 3, task & 7);
            colo[j] = (task & 7) * CHUNK;
        }
        auto set_rows = [&](int mt) {
#pragma unroll
            for (int j = 0; j < TASKS; ++j) {
                int64_t gm = (int64_t)mt * BM + ((ptid + j * NUM_A_THREADS) >> 3);
                if (gm >= M) gm = M - 1;                       // clamp: rows beyond M are never stored
                rowp[j] = A + gm * lda + colo[j];
            }
        };
        auto load_a = [&](int kb, float4* r) {
            const int k0 = kb * BK;
            if (vecA && (k0 + BK <= K)) {
#pragma unroll
                for (int j = 0; j < TASKS; ++j)
#pragma unroll
                    for (int q = 0; q < V; ++q) r[V * j + q] = __ldg(reinterpret_cast<const float4*>(rowp[j] + k0) + q);
            } else {
#pragma unroll
                for (int j = 0; j < TASKS; ++j) {
                    float v[CHUNK];
#pragma unroll
                    for (int e = 0; e < CHUNK; ++e) v[e] = (k0 + colo[j] + e < K) ? __ldg(rowp[j] + k0 + e) : 0.f;
#pragma unroll
                    for (int q = 0; q < V; ++q) r[V * j + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        // the load cursor (lmt, lnt, lkb) runs PF-1 k-blocks ahead of the convert cursor, across tile boundaries;
        // `ahead` = k-blocks loaded but not yet converted
        int lmt = mt0, lnt = nt0, lkb = 0;
        bool lvalid = lmt < m_tiles;
        if (lvalid) set_rows(lmt);
        auto advance = [&]() {
            if (++lkb == KB) { lkb = 0; tile_next(lmt, lnt); if (lmt < m_tiles) set_rows(lmt); }
            lvalid = lmt < m_tiles;
        };
        int ahead = 0;
#pragma unroll
        for (int s = 0; s < PF - 1; ++s)
            if (lvalid) { load_a(lkb, buf[s]); ++ahead; advance(); }
        bool done = (ahead == 0);
        while (!done) {
#pragma unroll
            for (int s = 0; s < PF; ++s) {
                if (ahead == 0) { done = true; break; }
                if (lvalid) { load_a(lkb, buf[(s + PF - 1) % PF]); ++ahead; advance(); }
                if (!mbar_wait(empty_bar(stage), phase ^ 1, vstatus)) { done = true; break; }
                unsigned char* sa_hi = smem + stage * cfg::STAGE_BYTES;
                unsigned char* sa_lo = sa_hi + cfg::A_BYTES;
#pragma unroll
                for (int j = 0; j < TASKS; ++j) {
                    uint4 hi, lo;
                    if constexpr (TF32) {
                        split4_tf32(buf[s][j], hi, lo);
                    } else {
                        const float v[8] = {buf[s][2 * j].x, buf[s][2 * j].y, buf[s][2 * j].z, buf[s][2 * j].w,
                                            buf[s][2 * j + 1].x, buf[s][2 * j + 1].y, buf[s][2 * j + 1].z, buf[s][2 * j + 1].w};
                        if constexpr (MODE == MODE_F16S) split8_f16s(v, hi, lo);
                        else split8(v, hi, lo);
                    }
                    *reinterpret_cast<uint4*>(sa_hi + soff[j]) = hi;
                    *reinterpret_cast<uint4*>(sa_lo + soff[j]) = lo;
                }
                if constexpr (!CF) fence_proxy_async();      // generic-proxy writes -> visible to the tensor core
                mbar_arrive(full_bar(stage));
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
                --ahead;
            }
        }
    } else if (warp >= B_WARP0) {
        // ================================================================ B loaders (pre-split, cp.async)
        // Two stages in flight: the copies of k-block i+1 are issued before waiting for those of k-block i
        // (cp.async groups), so the L2 latency of one stage overlaps the issue of the next.
        const int ptid = tid - B_WARP0 * 32;
        int stage = 0, prev_stage = -1;
        uint32_t phase = 0;
        bool ok = true;
        constexpr int BT = (BN * 8 + NUM_B_THREADS - 1) / NUM_B_THREADS;      // tasks per thread (last one guarded)
        const uint32_t row_bytes = (uint32_t)Kp * ESZ;
        for (int mt = mt0, nt = nt0; mt < m_tiles && ok; tile_next(mt, nt)) {
            const uint32_t tile_off = (uint32_t)(nt * BN) * row_bytes;          // 32-bit: B is at most a few MB
            for (int kb = 0; kb < KB; ++kb) {
                if (!mbar_wait(empty_bar(stage), phase ^ 1, vstatus)) { ok = false; break; }
                const uint32_t sb_hi = smem_u32(smem + stage * cfg::STAGE_BYTES + 2 * cfg::A_BYTES);
                const uint32_t sb_lo = sb_hi + cfg::B_BYTES;
                const uint32_t k_off = tile_off + (uint32_t)(kb * BK) * ESZ;
#pragma unroll
                for (int j = 0; j < BT; ++j) {
                    const int task = ptid + j * NUM_B_THREADS;
                    if (task >= BN * 8) break;
                    const int row = task >> 3, ch = task & 7;
                    const uint32_t e = k_off + (uint32_t)row * row_bytes + (uint32_t)(ch * CHUNK) * ESZ;
                    const uint32_t off = swz(row, ch);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_hi + off), "l"(Bhi + e) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_lo + off), "l"(Blo + e) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (prev_stage >= 0) {
                    asm volatile("cp.async.wait_group 1;" ::: "memory");      // the previous stage has landed
                    if constexpr (!CF) fence_proxy_async();
                    mbar_arrive(full_bar(prev_stage));
                }
                prev_stage = stage;
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (prev_stage >= 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if constexpr (!CF) fence_proxy_async();
            if (ok) mbar_arrive(full_bar(prev_stage));
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer
        constexpr uint32_t idesc = make_idesc(BM, BN, MODE);
        constexpr uint32_t idesc2 = make_idesc(BM, (TWOACC && BN <= 128) ? 2 * BN : BN, MODE);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        bool ok = true;
        for (int mt = mt0, nt = nt0; mt < m_tiles && ok; tile_next(mt, nt)) {
            if (!mbar_wait(tempty_bar(acc), acc_phase ^ 1, vstatus)) { ok = false; break; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * cfg::NACC * BN);
            const uint32_t tmem_x = TWOACC ? tmem_d + BN : tmem_d;        // cross-term accumulator
            for (int kb = 0; kb < KB; ++kb) {
                if (!mbar_wait(full_bar(stage), phase, vstatus)) { ok = false; break; }
                tc_fence_after();
                if (elect_one()) {
                    if constexpr (CF) fence_proxy_async();        // acquired the producers' writes: hand them to the async proxy
                    const uint32_t sa_hi = smem_u32(smem + stage * cfg::STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + cfg::A_BYTES;
                    const uint32_t sb_hi = sa_lo + cfg::A_BYTES;
                    const uint32_t sb_lo = sb_hi + cfg::B_BYTES;
                    const uint64_t dah = make_desc(sa_hi), dal = make_desc(sa_lo);
                    const uint64_t dbh = make_desc(sb_hi), dbl = make_desc(sb_lo);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {                      // 4 MMA k-steps of 32 bytes per row
                        const uint64_t adv = (uint64_t)(kk * 2);          // 32 bytes >> 4
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        if constexpr (TF32) {
                            // [main | cross] += Ahi x [Bhi ; Blo] as ONE N = 2*BN instruction (the two B halves and
                            // the two accumulators are adjacent), then cross += Alo x Bhi: the A tile is read from
                            // shared memory twice instead of three times
                            umma<true>(tmem_d, dah + adv, dbh + adv, idesc2, first);
                            umma<true>(tmem_x, dal + adv, dbh + adv, idesc, 1u);
                        } else if constexpr (MODE == MODE_F16S) {
                            // same two-instruction scheme at the kind::f16 rate; both cross terms carry the 2^11
                            // scale of the residuals
                            umma<false>(tmem_d, dah + adv, dbh + adv, idesc2, first);
                            umma<false>(tmem_x, dal + adv, dbh + adv, idesc, 1u);
                        } else {
                            umma<false>(tmem_d, dah + adv, dbh + adv, idesc, first);
                            umma<false>(tmem_d, dah + adv, dbl + adv, idesc, 1u);
                            umma<false>(tmem_d, dal + adv, dbh + adv, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));        // frees the stage once the MMAs have read it
                }
                __syncwarp();
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (!ok) break;
            if (elect_one()) umma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ================================================================ epilogue (warps 0..7)
        // TMEM gives each thread one accumulator row (a warp can only read the lane quarter warp % 4); 32x32
        // blocks go through a swizzled shared tile so that global stores are 128-byte contiguous per row
        // (8 lanes x float4).  Warps w and w + 4 share a lane quarter and alternate 32-column chunks: with K <= 128
        // the epilogue, not the MMAs, paces the kernel (measured: the MMA warp spent ~60 % of its time waiting for
        // a free accumulator with 4 epilogue warps).
        const int quarter = warp & 3, half = warp >> 2;
        float* T = epi_smem + warp * (32 * EPI_LD);
        int acc = 0;
        uint32_t acc_phase = 0;
        bool ok = true;
        for (int mt = mt0, nt = nt0; mt < m_tiles && ok; tile_next(mt, nt)) {
            const int64_t m0 = (int64_t)mt * BM + quarter * 32;
            const int n0 = nt * BN;
            if (!mbar_wait(tfull_bar(acc), acc_phase, vstatus)) { ok = false; break; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * cfg::NACC * BN);
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                float v[32];
                // this lane's 4 output columns are the same for all 8 row groups of the chunk: ONE bias load per
                // chunk, issued before the TMEM read (it used to be re-loaded in front of every store: eight
                // dependent global-load stalls per chunk, ~27 % of the epilogue warps' time on K <= 128 shapes)
                const int colv = n0 + c0 + (lane & 7) * 4;
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool bias_v = bias != nullptr && vecC && colv + 4 <= N;
                if (bias_v) bb = __ldg(reinterpret_cast<const float4*>(bias + colv));
                {
                    uint32_t rv[32], rw[32];
                    tmem_ld32_issue(taddr + c0, rv);    // all lanes participate (sync.aligned); one wait for both
                    if constexpr (TWOACC) tmem_ld32_issue(taddr + BN + c0, rw);   // cross-term accumulator
                    tmem_ld_wait();
                    tmem_pin32(rv);
                    if constexpr (TF32) {
                        tmem_pin32(rw);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rv[j]) + __uint_as_float(rw[j]);
                    } else if constexpr (MODE == MODE_F16S) {
                        tmem_pin32(rw);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = fmaf(__uint_as_float(rw[j]), 1.f / 2048.f, __uint_as_float(rv[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rv[j]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(T + lane * EPI_LD + 4 * (q ^ (lane & 7))) =
                        make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                __syncwarp();
                if (vecC && m0 + 32 <= M && n0 + c0 + 32 <= N) {
                    // interior chunk (every chunk of the step's shapes): no per-element bounds tests or index
                    // arithmetic -- one pointer walking 4 rows per step.  (The guarded loop below costs ~580 warp
                    // instructions per 32 x 32 chunk and made the epilogue warps the pace of every K <= 128 GEMM.)
                    const int g = lane & 7;
                    float* cp = C + (m0 + (lane >> 3)) * ldc + colv;
                    const int64_t step = 4 * ldc;
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = rr * 4 + (lane >> 3);
                        float4 o = *reinterpret_cast<const float4*>(T + r * EPI_LD + 4 * (g ^ (r & 7)));
                        if (bias_v) { o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
                        if (accumulate) {
                            const float4 old = *reinterpret_cast<const float4*>(cp);
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *reinterpret_cast<float4*>(cp) = o;
                        cp += step;
                    }
                } else if (n0 + c0 < N) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = rr * 4 + (lane >> 3), g = lane & 7, cq = g * 4;
                        const int64_t row = m0 + r;
                        const int col = n0 + c0 + cq;
                        if (row < M && col < N) {
                            float4 o = *reinterpret_cast<const float4*>(T + r * EPI_LD + 4 * (g ^ (r & 7)));
                            float* cp = C + row * ldc + col;
                            if (vecC && col + 4 <= N) {
                                if (bias_v) { o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
                                if (accumulate) {
                                    const float4 old = *reinterpret_cast<const float4*>(cp);
                                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                }
                                *reinterpret_cast<float4*>(cp) = o;
                            } else {
                                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (col + j < N) {
                                        float x = ov[j];
                                        if (bias) x += __ldg(bias + col + j);
                                        if (accumulate) x += cp[j];
                                        cp[j] = x;
                                    }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, cfg::TMEM_COLS);
    }
}

// B (weights) -> zero-padded hi / lo terms, K-major [Npad, Kp] (bf16 or tf32-in-fp32); clears the status word
template <int MODE>
__global__ void presplit_b_kernel(const float* __restrict__ B, int64_t ldb, int transB, int N, int K, int Npad, int Kp,
                                  void* __restrict__ hi_, void* __restrict__ lo_, int* status) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *status = 0;
    const int64_t total = (int64_t)Npad * Kp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / Kp), k = (int)(i % Kp);
        float v = 0.f;
        if (n < N && k < K) v = transB ? __ldg(B + (int64_t)n * ldb + k) : __ldg(B + (int64_t)k * ldb + n);
        if constexpr (MODE == MODE_TF32) {
            const uint32_t h = tf32_rna_bits(v);
            const uint32_t l = tf32_rna_bits(v - __uint_as_float(h));
            reinterpret_cast<uint32_t*>(hi_)[i] = h;
            reinterpret_cast<uint32_t*>(lo_)[i] = l;
        } else if constexpr (MODE == MODE_F16S) {
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
            reinterpret_cast<uint16_t*>(hi_)[i] = __half_as_ushort(h);
            reinterpret_cast<uint16_t*>(lo_)[i] = __half_as_ushort(l);
        } else {
            const uint32_t u = __float_as_uint(v);
            const float r = v - __uint_as_float(u & 0xffff0000u);
            uint32_t l2;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l2) : "f"(0.f), "f"(r));
            reinterpret_cast<uint16_t*>(hi_)[i] = (uint16_t)(u >> 16);
            reinterpret_cast<uint16_t*>(lo_)[i] = (uint16_t)(l2 & 0xffffu);
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <int BN, int MODE, int PF, bool CF>
int launch_tc_pf(int64_t M, int N, int K, const float* A, int64_t lda, const unsigned char* hi, const unsigned char* lo,
                 int Kp, int Npad, float* C, int64_t ldc, const float* bias, int accumulate, int* status, cudaStream_t st) {
    using cfg = Cfg<BN, MODE>;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, PF, CF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = Npad / BN;
    const int64_t total = ((M + BM - 1) / BM) * n_tiles;
    const int grid = (int)(total < kNumSMs ? total : kNumSMs);
    const bool vecA = (lda % 4 == 0) && aligned16(A);
    const bool vecC = (ldc % 4 == 0) && aligned16(C) && (bias == nullptr || aligned16(bias));
    gemm_tc_kernel<BN, MODE, PF, CF><<<grid, TC_THREADS, cfg::SMEM_BYTES, st>>>(M, N, K, A, lda, hi, lo, Kp, n_tiles, C, ldc,
                                                                             bias, accumulate, status, vecA, vecC);
    return spgan_launch_status();
}

// A-producer prefetch depth (register sets): SPGAN_TC_PF = 2 | 3 | 4 overrides the default (tuning knob)
inline int tc_prefetch_sets() {
    static const int pf = [] {
        const char* e = getenv("SPGAN_TC_PF");
        const int v = e ? atoi(e) : 0;
        return (v >= 2 && v <= 4) ? v : SPGAN_TC_PF_DEFAULT;
    }();
    return pf;
}

// SPGAN_TC_FENCE = producer | consumer overrides where the proxy fence runs (see gemm_tc_kernel)
inline bool tc_consumer_fence() {
    static const bool cf = [] {
        const char* e = getenv("SPGAN_TC_FENCE");
        if (e && e[0] == 'p') return false;
        if (e && e[0] == 'c') return true;
        return SPGAN_TC_CONSUMER_FENCE_DEFAULT != 0;
    }();
    return cf;
}

#define SPGAN_TC_ARGS M, N, K, A, lda, hi, lo, Kp, Npad, C, ldc, bias, accumulate, status, st
template <int BN, int MODE>
int launch_tc(int64_t M, int N, int K, const float* A, int64_t lda, const unsigned char* hi, const unsigned char* lo,
              int Kp, int Npad, float* C, int64_t ldc, const float* bias, int accumulate, int* status, cudaStream_t st) {
    const bool cf = tc_consumer_fence();
    if constexpr (MODE == MODE_TF32) {
        const int pf = tc_prefetch_sets();
        if (pf == 3) return cf ? launch_tc_pf<BN, MODE, 3, true>(SPGAN_TC_ARGS) : launch_tc_pf<BN, MODE, 3, false>(SPGAN_TC_ARGS);
        if (pf == 4) return cf ? launch_tc_pf<BN, MODE, 4, true>(SPGAN_TC_ARGS) : launch_tc_pf<BN, MODE, 4, false>(SPGAN_TC_ARGS);
    }
    // 2-byte modes: 8 floats per task, two register sets
    return cf ? launch_tc_pf<BN, MODE, 2, true>(SPGAN_TC_ARGS) : launch_tc_pf<BN, MODE, 2, false>(SPGAN_TC_ARGS);
}
#undef SPGAN_TC_ARGS

template <int MODE>
int run_tc(int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
           int64_t ldc, const float* bias, int accumulate, void* workspace, cudaStream_t st) {
    constexpr bool TF32 = MODE == MODE_TF32;
    constexpr int BK = TF32 ? 32 : 64;
    constexpr size_t ESZ = TF32 ? 4 : 2;
    const int BN = N <= 64 ? 64 : ((N <= 128 || MODE != MODE_BF16) ? 128 : 256);     // two accumulators per tile: BN <= 128
    const int Npad = (int)align_up((size_t)N, BN), Kp = (int)align_up((size_t)K, BK);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int* status = reinterpret_cast<int*>(ws);
    unsigned char* hi = ws + 256;
    unsigned char* lo = hi + align_up(align_up((size_t)N, 256) * Kp * ESZ, 256);
    int rc = SPGAN_OK;
    if ((transB & 2) == 0) {             // bit 1 of transB: the workspace already holds this weight's split operand
        presplit_b_kernel<MODE><<<ew_grid((int64_t)Npad * Kp, 256), 256, 0, st>>>(B, ldb, transB & 1, N, K, Npad, Kp, hi, lo, status);
        rc = spgan_launch_status();
        if (rc != SPGAN_OK) return rc;
    }
    if (BN == 64) return launch_tc<64, MODE>(M, N, K, A, lda, hi, lo, Kp, Npad, C, ldc, bias, accumulate, status, st);
    if (BN == 128) return launch_tc<128, MODE>(M, N, K, A, lda, hi, lo, Kp, Npad, C, ldc, bias, accumulate, status, st);
    if constexpr (MODE == MODE_BF16)
        return launch_tc<256, MODE_BF16>(M, N, K, A, lda, hi, lo, Kp, Npad, C, ldc, bias, accumulate, status, st);
    return SPGAN_E_UNSUPPORTED;
}

}  // namespace

// workspace big enough for either mode: status word + hi + lo, each up to roundup(N,256) x roundup(K,64) fp32
size_t spgan_gemm_tc_workspace(int N, int K) {
    const size_t Npad = align_up((size_t)N, 256), Kp = align_up((size_t)K, 64);
    return 256 + 2 * align_up(Npad * Kp * sizeof(float), 256);
}

bool spgan_gemm_tc_supported(int transA, int64_t M, int N, int K) {
    return !transA && M >= 128 && N >= 16 && K >= 16;
}

// mode: 0 = TF32x3 (engine 1), 1 = BF16x3 (engine 2), 2 = FP16Sx3 (engine 3)
int spgan_gemm_tc(int mode, int transB, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, float* C, int64_t ldc, const float* bias, int accumulate, void* workspace,
                  cudaStream_t st) {
    if (mode == MODE_BF16) return run_tc<MODE_BF16>(transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, workspace, st);
    if (mode == MODE_F16S) return run_tc<MODE_F16S>(transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, workspace, st);
    return run_tc<MODE_TF32>(transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, workspace, st);
}
