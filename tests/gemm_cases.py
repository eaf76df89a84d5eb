"""Shared GEMM case runner for tests/test_gemm_gpu.py and scripts/gemm_debug.py."""
import torch

from mtrl_b200 import _lib as L


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 emulation: round to nearest, ties away from zero, 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    r = (i + 0x1000) & ~0x1FFF
    return r.view(torch.float32)


def make_problem(M, N, K, a_major, b_major, epilogue, block_n=256, k_splits=1, seed=0, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = tf32_round(torch.randn(M, K, generator=g)).to(dev)   # logical A[m][k]
    B = tf32_round(torch.randn(N, K, generator=g)).to(dev)   # logical B[n][k]
    bias = torch.randn(N, generator=g).to(dev)
    mask = torch.randn(M, N, generator=g).to(dev)
    A_store = A.t().contiguous() if a_major else A.contiguous()
    B_store = B.t().contiguous() if b_major else B.contiguous()
    D = torch.zeros(M, N, device=dev)
    ref = A.double() @ B.double().t()
    if epilogue == L.EPI_BIAS_RELU:
        ref = torch.relu(ref + bias.double())
    elif epilogue == L.EPI_RELU_MASK:
        ref = ref * (mask > 0)
    p = L.GemmProblem(
        A=A_store.data_ptr(), lda=A_store.stride(0), a_major=a_major,
        B=B_store.data_ptr(), ldb=B_store.stride(0), b_major=b_major,
        D=D.data_ptr(), ldd=D.stride(0), M=M, N=N, K=K,
        block_n=block_n, k_splits=k_splits, epilogue=epilogue,
        bias=bias.data_ptr(), mask=mask.data_ptr(), ldmask=mask.stride(0),
    )
    keep = (A_store, B_store, bias, mask)
    return p, D, ref, keep


def split_tf32(x: torch.Tensor):
    """(hi, lo) with hi = tf32(x), lo = tf32(x - hi): the operand pair of the fp32x3 mode."""
    hi = tf32_round(x)
    return hi, tf32_round(x - hi)


def make_problem_x3(M, N, K, a_major, b_major, epilogue, block_n=256, k_splits=1, seed=0, dev="cuda", with_d_lo=True):
    """Same contraction on FULL fp32 operands given as (hi, lo) tf32 pairs; the reference is the fp64 product of the
    unsplit fp32 values.  Rounding epilogues also emit D_lo (returned) so that D + D_lo is the unrounded result."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    mask = torch.randn(M, N, generator=g).to(dev)
    lay = lambda X, major: X.t().contiguous() if major else X.contiguous()  # noqa: E731
    (Ah, Al), (Bh, Bl) = split_tf32(A), split_tf32(B)
    Ah, Al, Bh, Bl = lay(Ah, a_major), lay(Al, a_major), lay(Bh, b_major), lay(Bl, b_major)
    D = torch.zeros(M, N, device=dev)
    rounds = epilogue in (L.EPI_BIAS_RELU, L.EPI_RELU_MASK, L.EPI_STORE_TF32)
    D_lo = torch.zeros(M, N, device=dev) if (rounds and with_d_lo) else None
    ref = A.double() @ B.double().t()
    if epilogue == L.EPI_BIAS_RELU:
        ref = torch.relu(ref + bias.double())
    elif epilogue == L.EPI_RELU_MASK:
        ref = ref * (mask > 0)
    p = L.GemmProblem(
        A=Ah.data_ptr(), lda=Ah.stride(0), a_major=a_major, B=Bh.data_ptr(), ldb=Bh.stride(0), b_major=b_major,
        D=D.data_ptr(), ldd=D.stride(0), M=M, N=N, K=K, block_n=block_n, k_splits=k_splits, epilogue=epilogue,
        bias=bias.data_ptr(), mask=mask.data_ptr(), ldmask=mask.stride(0),
        A_lo=Al.data_ptr(), B_lo=Bl.data_ptr(), D_lo=D_lo.data_ptr() if D_lo is not None else None,
    )
    return p, D, D_lo, ref, (Ah, Al, Bh, Bl, bias, mask)


def run_case_x3(case, dev="cuda", with_d_lo=True, ctas=0):
    name, M, N, K, am, bm, epi, bn, ks = case
    p, D, D_lo, ref, keep = make_problem_x3(M, N, K, am, bm, epi, bn, ks, dev=dev, with_d_lo=with_d_lo)
    L.GemmPlan([p], ctas=ctas).run()
    torch.cuda.synchronize()
    out = D.double() + (D_lo.double() if D_lo is not None else 0)
    return ((out - ref).norm() / ref.norm().clamp_min(1e-30)).item(), D, D_lo


def rel_err(D, ref, epilogue):
    out = D.double()
    if epilogue in (L.EPI_BIAS_RELU, L.EPI_RELU_MASK, L.EPI_STORE_TF32):
        # outputs are rounded to tf32 (2^-11 relative) on top of fp32 accumulation
        pass
    return ((out - ref).norm() / ref.norm().clamp_min(1e-30)).item(), (out - ref).abs().max().item()


CASES = [
    # name, M, N, K, a_major, b_major, epilogue, block_n, k_splits
    ("kk_store_small", 128, 256, 32, 0, 0, L.EPI_STORE, 256, 1),
    ("kk_store", 256, 512, 256, 0, 0, L.EPI_STORE, 256, 1),
    ("k_mn_fwd_small", 128, 256, 32, 0, 1, L.EPI_STORE, 256, 1),
    ("k_mn_fwd_biasrelu", 384, 512, 320, 0, 1, L.EPI_BIAS_RELU, 256, 1),
    ("mn_k_small", 128, 256, 32, 1, 0, L.EPI_STORE, 256, 1),
    ("mn_mn_dw_small", 128, 256, 32, 1, 1, L.EPI_STORE, 256, 1),
    ("mn_mn_dw", 512, 512, 640, 1, 1, L.EPI_STORE, 256, 1),
    ("mn_mn_dw_splitk", 512, 512, 1280, 1, 1, L.EPI_ATOMIC_ADD, 256, 3),
    ("kk_dx_mask", 384, 512, 512, 0, 0, L.EPI_RELU_MASK, 256, 1),
    ("ragged_w400_fwd", 256, 400, 400, 0, 1, L.EPI_BIAS_RELU, 208, 1),
    ("ragged_w400_dx", 256, 400, 400, 0, 0, L.EPI_RELU_MASK, 208, 1),
    ("ragged_w400_dw", 400, 400, 256, 1, 1, L.EPI_STORE, 208, 1),
    ("layer0_fwd_k96", 256, 512, 96, 0, 1, L.EPI_BIAS_RELU, 256, 1),
    ("layer0_dw_m93", 93 + 3, 512, 256, 1, 1, L.EPI_ATOMIC_ADD, 256, 2),
    ("dx0_n16", 256, 16, 512, 0, 0, L.EPI_STORE_TF32, 16, 1),
    ("n64_blockn64", 256, 64, 128, 0, 1, L.EPI_STORE, 64, 1),
]


def run_case(case, dev="cuda", ctas=0):
    name, M, N, K, am, bm, epi, bn, ks = case
    p, D, ref, keep = make_problem(M, N, K, am, bm, epi, bn, ks, dev=dev)
    plan = L.GemmPlan([p], ctas=ctas)
    assert (ctas & 15) == 0 or plan.ctas == (ctas & 15)
    plan.run()
    torch.cuda.synchronize()
    r, a = rel_err(D, ref, epi)
    return r, a


def run_grouped(dev="cuda"):
    specs = [CASES[1], CASES[3], CASES[6], CASES[8]]
    probs, outs = [], []
    for i, (name, M, N, K, am, bm, epi, bn, ks) in enumerate(specs):
        p, D, ref, keep = make_problem(M, N, K, am, bm, epi, bn, ks, seed=10 + i, dev=dev)
        probs.append(p)
        outs.append((name, D, ref, epi, keep))
    plan = L.GemmPlan(probs)
    plan.run()
    torch.cuda.synchronize()
    return [(name,) + rel_err(D, ref, epi) for name, D, ref, epi, _ in outs]
