"""tcgen05 grouped GEMM vs an fp64 matmul of the same tf32-rounded operands."""
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    import gemm_cases as G

    return G.CASES


@pytest.mark.parametrize("idx", range(16))
def test_gemm_case(cuda, idx):
    import gemm_cases as G

    case = G.CASES[idx]
    rel, _ = G.run_case(case)
    # fp32 accumulation of exact tf32 products; tf32-rounded outputs add <= 2^-11 relative
    assert rel < 5e-4, f"{case[0]}: rel err {rel}"


def test_gemm_grouped(cuda):
    import gemm_cases as G

    for name, rel, _ in G.run_grouped():
        assert rel < 5e-4, f"{name}: rel err {rel}"


def test_relu_mask_epilogue_emits_column_sum_partials(cuda):
    """The dX epilogue also writes, per 32-row group, the column sums of what it stored (bias-gradient partials)."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, N, K = 384, 400, 160
    p, D, ref, keep = G.make_problem(M, N, K, 0, 0, L.EPI_RELU_MASK, 208, 1, seed=3)
    part = torch.full(((M + 31) // 32, N), float("nan"), device="cuda")
    p.colsum_partial = part.data_ptr()
    plan = L.GemmPlan([p])
    plan.run()
    torch.cuda.synchronize()
    expect = D.double().reshape(M // 32, 32, N).sum(1)
    assert torch.isfinite(part).all()
    assert ((part.double() - expect).norm() / expect.norm()).item() < 1e-6


def test_relu_bits_roundtrip(cuda):
    """Forward epilogue emits the ReLU sign bits; the dX epilogue consuming those bits equals the float-mask version."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, N, K = 384, 448, 192
    pf, H, ref_h, keep_f = G.make_problem(M, N, K, 0, 1, L.EPI_BIAS_RELU, 224, 1, seed=5)
    bits = torch.zeros(M, (N + 31) // 32, dtype=torch.int32, device="cuda")
    pf.relu_bits_out = bits.data_ptr()
    pf.ldbits = bits.shape[1]
    L.GemmPlan([pf]).run()
    torch.cuda.synchronize()
    expect = (H > 0)
    got = ((bits.unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(M, -1)[:, :N].bool()
    assert torch.equal(got, expect)
    # dX with bits vs dX with the float activation as mask
    pb, D1, _, keep1 = G.make_problem(M, N, 256, 0, 0, L.EPI_RELU_MASK, 224, 1, seed=6)
    pb.mask = None
    pb.mask_bits = bits.data_ptr()
    pb.ldbits = bits.shape[1]
    pm, D2, _, keep2 = G.make_problem(M, N, 256, 0, 0, L.EPI_RELU_MASK, 224, 1, seed=6)
    pm.mask = H.data_ptr()
    pm.ldmask = H.stride(0)
    L.GemmPlan([pb]).run()
    L.GemmPlan([pm]).run()
    torch.cuda.synchronize()
    assert torch.equal(D1, D2)
