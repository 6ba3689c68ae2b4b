// tcgen05 weight-gradient GEMM: C[Mo,No] (+)= A^T B with A = dY [K, Mo], B = X [K, No] row-major fp32,
// K = number of points / edges (1e5 .. 1e6), Mo/No = channel counts.  TF32x3 split like gemm_tc.cu.
//
// Both operands are contiguous along their M / N dimension while the tensor core wants K-major tiles.
// The producers transpose on the fly: lane = output channel, each thread gathers 4 consecutive k for its
// channel with four warp-coalesced 128-byte loads and writes them as ONE 16-byte chunk of the channel's
// 128-byte swizzled K-major row (conflict-free), so the proven K-major descriptors of gemm_tc.cu apply.
//
// Work unit = (output tile 128 x BN, chunk of KC = 1024 rows).  A unit accumulates 32 k-blocks in TMEM and
// is flushed with fp32 vector atomics into C, which bounds the truncating tensor-core accumulation chain
// (error grows with chain length) and provides the split-K needed to fill 148 SMs with <= 80 output tiles.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 32;                       // tf32 elements per k-block
constexpr int KC = 1024;                     // rows per work unit
constexpr int NUM_EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int P_WARP0 = 5;
constexpr int NUM_P_THREADS = 16 * 32;       // 16 producer warps: scalar transposing loads need the parallelism
constexpr int TN_THREADS = 21 * 32;          // 672
constexpr int EPI_LD = 36;

template <int BN>
struct CfgTN {
    static constexpr int A_BYTES = BK * BM * 4;           // one half: 16 KB
    static constexpr int B_BYTES = BK * BN * 4;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int TMEM_COLS = 2 * BN;              // main + cross-term accumulator, one buffer
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + NUM_EPI_WARPS * 32 * EPI_LD * 4;
};

// K-major, 128-byte swizzle (same as gemm_tc.cu): rows of 32 tf32, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D=f32, A=B=tf32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_k(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int BN>
__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tc_tn_kernel(int Mo, int No, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                  int64_t ldb, float* __restrict__ C, int64_t ldc, int m_tiles, int n_tiles, int64_t k_chunks,
                  int* status, bool vecC) {
    using cfg = CfgTN<BN>;
    constexpr int A_TASKS = BM * (BK / 4) / NUM_P_THREADS;     // 2 (row, 4-k chunk) tasks per thread
    constexpr int B_TASKS = BN * (BK / 4) / NUM_P_THREADS;     // 1 / 2 / 4
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = tc::align_smem_1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * cfg::STAGES + 2);
    float* epi_smem = reinterpret_cast<float*>(smem + cfg::STAGES * cfg::STAGE_BYTES + 256);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (see tc::elect_one)
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (cfg::STAGES + s); };
    const uint32_t tfull_bar = bar0 + 8u * (2 * cfg::STAGES);
    const uint32_t tempty_bar = bar0 + 8u * (2 * cfg::STAGES + 1);

    if (tid == 0) {
        for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(full_bar(s), NUM_P_THREADS); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, NUM_EPI_WARPS * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int64_t tiles = (int64_t)m_tiles * n_tiles;
    const int64_t units = tiles * k_chunks;
    volatile int* vstatus = status;
    // unit u -> tile (u % tiles), chunk (u / tiles): CTAs running side by side share the chunk's B rows in L2
    auto unit_coords = [&](int64_t u, int& m0, int& n0, int64_t& kbeg, int& kblocks) {
        const int64_t tile = u % tiles, chunk = u / tiles;
        m0 = (int)(tile / n_tiles) * BM;
        n0 = (int)(tile % n_tiles) * BN;
        kbeg = chunk * KC;
        const int64_t kend = kbeg + KC < K ? kbeg + KC : K;
        kblocks = (int)((kend - kbeg + BK - 1) / BK);
    };

    if (warp >= P_WARP0) {
        // ================================================================ producers: both operands
        const int ptid = tid - P_WARP0 * 32;
        float4 cur[A_TASKS + B_TASKS], nxt[A_TASKS + B_TASKS];
        // task -> (channel row, 16-byte chunk = 4 consecutive k): lanes walk channels => coalesced rows
        // loop invariants per task: swizzled shared offsets and the (row, first k) of the chunk
        uint32_t soffA[A_TASKS], soffB[B_TASKS];
#pragma unroll
        for (int j = 0; j < A_TASKS; ++j) { const int task = ptid + j * NUM_P_THREADS; soffA[j] = swz(task & (BM - 1), task / BM); }
#pragma unroll
        for (int j = 0; j < B_TASKS; ++j) { const int task = ptid + j * NUM_P_THREADS; soffB[j] = swz(task % BN, task / BN); }
        auto load_kb = [&](int m0, int n0, int64_t k0, float4* r) {
            const bool fullk = (k0 + BK <= K);                       // whole k-block in range: no per-element guards
#pragma unroll
            for (int j = 0; j < A_TASKS; ++j) {
                const int task = ptid + j * NUM_P_THREADS;
                const int m = m0 + (task & (BM - 1));
                const int64_t k = k0 + (task / BM) * 4;
                const float* p = A + k * lda + m;
                float v[4];
                if (fullk && m < Mo) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = __ldg(p + e * lda);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = (m < Mo && k + e < K) ? __ldg(p + e * lda) : 0.f;
                }
                r[j] = make_float4(v[0], v[1], v[2], v[3]);
            }
#pragma unroll
            for (int j = 0; j < B_TASKS; ++j) {
                const int task = ptid + j * NUM_P_THREADS;
                const int n = n0 + (task % BN);
                const int64_t k = k0 + (task / BN) * 4;
                const float* p = B + k * ldb + n;
                float v[4];
                if (fullk && n < No) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = __ldg(p + e * ldb);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = (n < No && k + e < K) ? __ldg(p + e * ldb) : 0.f;
                }
                r[A_TASKS + j] = make_float4(v[0], v[1], v[2], v[3]);
            }
        };
        int stage = 0;
        uint32_t phase = 0;
        int64_t u = blockIdx.x;
        int m0 = 0, n0 = 0, kblocks = 0, kb = 0;
        int64_t kbeg = 0;
        bool have = u < units;
        if (have) { unit_coords(u, m0, n0, kbeg, kblocks); load_kb(m0, n0, kbeg, cur); }
        while (have) {
            // coordinates of the next k-block (possibly of the next unit)
            int64_t un = u; int m0n = m0, n0n = n0, kblocksn = kblocks, kbn = kb + 1; int64_t kbegn = kbeg;
            if (kbn == kblocks) { un += gridDim.x; kbn = 0; if (un < units) unit_coords(un, m0n, n0n, kbegn, kblocksn); }
            const bool have_next = un < units;
            if (have_next) load_kb(m0n, n0n, kbegn + (int64_t)kbn * BK, nxt);
            if (!mbar_wait(empty_bar(stage), phase ^ 1, vstatus)) break;
            unsigned char* sa_hi = smem + stage * cfg::STAGE_BYTES;
            unsigned char* sa_lo = sa_hi + cfg::A_BYTES;
            unsigned char* sb_hi = sa_lo + cfg::A_BYTES;
            unsigned char* sb_lo = sb_hi + cfg::B_BYTES;
#pragma unroll
            for (int j = 0; j < A_TASKS; ++j) {
                const int task = ptid + j * NUM_P_THREADS;
                uint4 hi, lo;
                split4_tf32(cur[j], hi, lo);
                *reinterpret_cast<uint4*>(sa_hi + soffA[j]) = hi;
                *reinterpret_cast<uint4*>(sa_lo + soffA[j]) = lo;
            }
#pragma unroll
            for (int j = 0; j < B_TASKS; ++j) {
                const int task = ptid + j * NUM_P_THREADS;
                uint4 hi, lo;
                split4_tf32(cur[A_TASKS + j], hi, lo);
                *reinterpret_cast<uint4*>(sb_hi + soffB[j]) = hi;
                *reinterpret_cast<uint4*>(sb_lo + soffB[j]) = lo;
            }
            fence_proxy_async();
            mbar_arrive(full_bar(stage));
            if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
            for (int j = 0; j < A_TASKS + B_TASKS; ++j) cur[j] = nxt[j];
            u = un; m0 = m0n; n0 = n0n; kbeg = kbegn; kblocks = kblocksn; kb = kbn; have = have_next;
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer
        constexpr uint32_t idesc = make_idesc_k(BM, BN);
        constexpr uint32_t idesc2 = make_idesc_k(BM, BN <= 128 ? 2 * BN : BN);
        int stage = 0;
        uint32_t phase = 0, tphase = 0;
        bool ok = true;
        const uint32_t tmem_d = tmem_base, tmem_x = tmem_base + BN;
        for (int64_t u = blockIdx.x; u < units && ok; u += gridDim.x) {
            int m0, n0, kblocks; int64_t kbeg;
            unit_coords(u, m0, n0, kbeg, kblocks);
            if (!mbar_wait(tempty_bar, tphase ^ 1, vstatus)) { ok = false; break; }
            tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb) {
                if (!mbar_wait(full_bar(stage), phase, vstatus)) { ok = false; break; }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa_hi = smem_u32(smem + stage * cfg::STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + cfg::A_BYTES;
                    const uint32_t sb_hi = sa_lo + cfg::A_BYTES;
                    const uint32_t sb_lo = sb_hi + cfg::B_BYTES;
                    const uint64_t dah = make_desc_k(sa_hi), dal = make_desc_k(sa_lo);
                    const uint64_t dbh = make_desc_k(sb_hi), dbl = make_desc_k(sb_lo);
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {               // 32 bytes per K=8 step
                        const uint64_t adv = (uint64_t)(kk * 2);
                        const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
                        if constexpr (BN <= 128) {
                            // [main | cross] += Ahi x [Bhi ; Blo] in one N = 2*BN instruction, then cross += Alo x Bhi
                            umma<true>(tmem_d, dah + adv, dbh + adv, idesc2, first);
                            umma<true>(tmem_x, dal + adv, dbh + adv, idesc, 1u);
                        } else {
                            umma<true>(tmem_d, dah + adv, dbh + adv, idesc, first);
                            umma<true>(tmem_x, dah + adv, dbl + adv, idesc, first);
                            umma<true>(tmem_x, dal + adv, dbh + adv, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (!ok) break;
            if (elect_one()) umma_commit(tfull_bar);
            __syncwarp();
            tphase ^= 1;
        }
    } else {
        // ================================================================ epilogue: atomic flush of the unit
        float* T = epi_smem + warp * (32 * EPI_LD);
        uint32_t tphase = 0;
        bool ok = true;
        for (int64_t u = blockIdx.x; u < units && ok; u += gridDim.x) {
            int m0, n0, kblocks; int64_t kbeg;
            unit_coords(u, m0, n0, kbeg, kblocks);
            if (!mbar_wait(tfull_bar, tphase, vstatus)) { ok = false; break; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                {
                    uint32_t rv[32], rw[32];
                    tmem_ld32_issue(taddr + c0, rv);              // main and cross-term accumulators: one wait
                    tmem_ld32_issue(taddr + BN + c0, rw);
                    tmem_ld_wait();
                    tmem_pin32(rv);
                    tmem_pin32(rw);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rv[j]) + __uint_as_float(rw[j]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(T + lane * EPI_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                __syncwarp();
                if (n0 + c0 < No) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = rr * 4 + (lane >> 3), cq = (lane & 7) * 4;
                        const int row = m0 + warp * 32 + r, col = n0 + c0 + cq;
                        if (row < Mo && col < No) {
                            const float4 o = *reinterpret_cast<const float4*>(T + r * EPI_LD + cq);
                            float* cp = C + (int64_t)row * ldc + col;
                            if (vecC && col + 4 <= No) {
                                atomicAdd(reinterpret_cast<float4*>(cp), o);
                            } else {
                                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (col + j < No) atomicAdd(cp + j, ov[j]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(tempty_bar);
            tphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, cfg::TMEM_COLS);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN>
int launch_tn(int Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
              int64_t ldc, int* status, cudaStream_t st) {
    using cfg = CfgTN<BN>;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int m_tiles = (Mo + BM - 1) / BM, n_tiles = (No + BN - 1) / BN;
    const int64_t k_chunks = (K + KC - 1) / KC;
    const int64_t units = (int64_t)m_tiles * n_tiles * k_chunks;
    const int grid = (int)(units < kNumSMs ? units : kNumSMs);
    const bool vecC = (ldc % 4 == 0) && aligned16(C);
    gemm_tc_tn_kernel<BN><<<grid, TN_THREADS, cfg::SMEM_BYTES, st>>>(Mo, No, K, A, lda, B, ldb, C, ldc, m_tiles,
                                                                      n_tiles, k_chunks, status, vecC);
    return spgan_launch_status();
}

}  // namespace

bool spgan_gemm_tc_tn_supported(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B,
                                int64_t ldb) {
    (void)A; (void)B; (void)lda; (void)ldb;          // scalar (warp-coalesced) operand loads: no alignment needs
    // measured: wins over the split-K CUDA-core kernel once the output tile work is large enough
    return Mo >= 16 && Mo <= 65536 && No >= 16 && K >= 4096 && (int64_t)Mo * No >= 8192;
}

// C[Mo,No] (+)= A^T B.  workspace: >= 256 bytes, zero-initialised by the caller (status word).
int spgan_gemm_tc_tn(int64_t Mo, int No, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                     int64_t ldc, int accumulate, void* workspace, cudaStream_t st) {
    // status word: written only by a pipeline timeout, which also traps (tc_common.cuh); the caller keeps it zeroed
    int* status = reinterpret_cast<int*>(workspace);
    if (!accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)No * sizeof(float), (size_t)Mo, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (No <= 64) return launch_tn<64>((int)Mo, No, K, A, lda, B, ldb, C, ldc, status, st);
    if (No <= 128) return launch_tn<128>((int)Mo, No, K, A, lda, B, ldb, C, ldc, status, st);
    return launch_tn<256>((int)Mo, No, K, A, lda, B, ldb, C, ldc, status, st);
}
