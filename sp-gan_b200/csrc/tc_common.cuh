// PTX wrappers shared by the tcgen05 GEMM kernels (mbarrier, TMEM, tcgen05.mma/ld/commit, proxy fences).
#pragma once
#include "common.cuh"

namespace tc {

constexpr uint32_t SPIN_LIMIT = 1u << 24;      // polls (>= ~60 cycles each): about a second before the guard trips

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 1024-byte alignment of the dynamic shared-memory window WITHOUT leaving the shared address space: offsetting the
// __shared__ array keeps every derived pointer a shared-space pointer (LDS / STS).  Rounding the generic address up
// through uintptr_t made all of them generic (LD.E / ST.E: long-scoreboard latency, and no reordering across the
// epilogue's global stores because the compiler had to assume aliasing).
__device__ __forceinline__ unsigned char* align_smem_1024(unsigned char* raw) {
    return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait (deadlock guard) as ONE asm block: a spin on mbarrier.try_wait with a poll counter.  A timeout is a
// protocol bug, never a load condition (the limit counts polls of this warp, not wall-clock time): the status word
// is raised for post-mortem reading and the kernel TRAPS, so the failure surfaces as a CUDA error at the caller's next
// synchronisation instead of a half-written output tensor flowing on into the step.
// The function never returns false.  Keeping the loop (and its exit) inside the asm block matters for code
// generation: a C-level `if (!wait(...)) return;` makes every loop-carried value of the calling role loop
// control-dependent on a per-thread predicate, and ptxas then treats the operands of UTCHMMA / UTMALDG as
// divergent (ELECT / R2UR.BROADCAST loops around every instruction).
// The exit is made WARP-UNIFORM by a vote: every caller is a converged warp, the barrier address and parity are
// warp-uniform, and once they live in uniform registers ptxas treats try_wait's predicate as uniform too -- it drops the
// reconvergence point after the loop and keeps the poll counter in a uniform register.  The hardware does hand
// different lanes different answers when the phase flips during the poll: lanes that left early then ran a
// .sync.aligned tcgen05.ld with a partial warp (measured: the tensor-core kNN filter lost neighbours for a few lanes of
// one warp, ~1 tile in 700).  With the vote all 32 lanes leave on the same poll.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* status) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 c;\n\t"
        "mov.u32 c, 0;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "vote.sync.all.pred p, p, 0xffffffff;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, %3;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "st.volatile.global.u32 [%2], 1;\n\t"
        "membar.sys;\n\t"
        "trap;\n\t"
        "WAIT_DONE:\n\t"
        "}"
        ::"r"(bar), "r"(parity), "l"(status), "r"(SPIN_LIMIT)
        : "memory");
    return true;
}
// One lane of a CONVERGED warp (elect.sync).  ptxas recognises the pattern: inside `if (elect_one())`, operands that
// are warp-uniform (kernel parameters, loop counters of warp-uniform loops, the warp index obtained through
// __shfl_sync) stay in uniform registers and UTCHMMA / UTMALDG issue directly.  Under a plain `if (lane == 0)` every
// operand is a per-thread value and each UTCHMMA is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~160
// cycles per instruction -- measured: the issuing warp, not the tensor pipe, paced the kernel).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::tf32 (K=8 per instruction) or kind::f16 (bf16, K=16)
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
    if constexpr (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// Batched TMEM reads: issue any number of tcgen05.ld, then ONE tcgen05.wait::ld.  (tmem_ld16 above waits after
// every 16 columns: four serialised TMEM round trips per 32-column chunk made the epilogue, not the MMAs, the
// bound of every K <= 128 GEMM.)  tmem_pin32 ties the loaded registers to a point after the wait so that the
// compiler cannot schedule their first use above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_pin32(uint32_t* r) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// fp32 -> tf32 with round-to-nearest (ties away from zero, like cvt.rna.tf32.f32) in two full-rate integer
// ops.  cvt.rna runs on the 16-lane/clk conversion unit and made the A-operand producers the bottleneck
// of the TF32x3 GEMM (2 conversions per element); IADD + LOP3 run at 128 lanes/clk.
__device__ __forceinline__ uint32_t tf32_rna_bits(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }

// 4 fp32 -> 4 tf32 "hi" (round to nearest) + 4 tf32 "lo" (rn of the exact remainder)
__device__ __forceinline__ void split4_tf32(float4 x, uint4& hi, uint4& lo) {
    const float v[4] = {x.x, x.y, x.z, x.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = tf32_rna_bits(v[i]);
        l[i] = tf32_rna_bits(v[i] - __uint_as_float(h[i]));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tc
