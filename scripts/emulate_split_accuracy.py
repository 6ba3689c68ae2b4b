"""CPU emulation (numpy) of the operand splits of the tcgen05 GEMM engines: rms / max relative error of
A.B^T against fp64 for  tf32x3 (engine 1),  bf16x3 (engine 2)  and  fp16x3 with 2^11-scaled residuals (engine 3),
with the cross terms summed separately and exact fp32-ish accumulation (float64 accumulate, rounded once: isolates
the SPLIT error from the tensor core's accumulation order).  No GPU needed:
    python scripts/emulate_split_accuracy.py"""
import numpy as np


def tf32(x):
    u = (x.astype(np.float32).view(np.uint32).astype(np.uint64) + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def bf16_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & 0xFFFF0000).view(np.float32)


def bf16_rn(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def f16(x):
    return x.astype(np.float16).astype(np.float32)


def product(Ah, Al, Bh, Bl, cross_scale=1.0):
    main = Ah.astype(np.float64) @ Bh.astype(np.float64).T
    cross = Ah.astype(np.float64) @ Bl.astype(np.float64).T + Al.astype(np.float64) @ Bh.astype(np.float64).T
    return (main.astype(np.float32) + (cross * cross_scale).astype(np.float32)).astype(np.float64)


def run(M, N, K, sa, sb, rng):
    A = (sa * rng.standard_normal((M, K))).astype(np.float32)
    B = (sb * rng.standard_normal((N, K))).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = {}
    Ah, Bh = tf32(A), tf32(B)
    out["tf32x3"] = product(Ah, tf32(A - Ah), Bh, tf32(B - Bh))
    Ah, Bh = bf16_trunc(A), bf16_trunc(B)
    out["bf16x3"] = product(Ah, bf16_rn(A - Ah), Bh, bf16_rn(B - Bh))
    Ah, Bh = f16(A), f16(B)
    out["fp16x3 scaled"] = product(Ah, f16((A - Ah) * 2048.0), Bh, f16((B - Bh) * 2048.0), 1.0 / 2048.0)
    out["fp16x3 unscaled"] = product(Ah, f16(A - Ah), Bh, f16(B - Bh))
    out["fp32 (rounded once)"] = ref.astype(np.float32).astype(np.float64)
    scale = np.sqrt((ref ** 2).mean())
    return {k: (np.sqrt(((v - ref) ** 2).mean()) / scale, np.abs(v - ref).max() / np.abs(ref).max()) for k, v in out.items()}


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for (M, N, K, sa, sb) in [(512, 256, 64, 1, 1), (512, 256, 256, 1, 1), (512, 128, 1280, 1, 1),
                              (512, 256, 256, 1e-3, 1e-2), (512, 256, 256, 300.0, 1e-3)]:
        r = run(M, N, K, sa, sb, rng)
        print("M=%d N=%d K=%d |A|~%g |B|~%g" % (M, N, K, sa, sb))
        for k, (rms, mx) in r.items():
            print("    %-22s rms %.2e   max %.2e" % (k, rms, mx))
