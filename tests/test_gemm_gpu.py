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


@pytest.mark.parametrize("idx", range(16))
def test_gemm_case_single_cta_tiles(cuda, idx):
    """The 128-row single-CTA tile shape (mtrl_gemm_plan_create_ex ctas = 1), which the plan autotuner picks for launches
    with too few 256-row units (one rank's rows of a sharded batch)."""
    import gemm_cases as G

    case = G.CASES[idx]
    rel, _ = G.run_case(case, ctas=1)
    assert rel < 5e-4, f"{case[0]}: rel err {rel}"
    if idx in (3, 6, 8):
        rel3, _, _ = G.run_case_x3(case, ctas=1)
        assert rel3 < 3e-6, f"{case[0]} fp32x3: rel err {rel3}"


@pytest.mark.parametrize("idx", range(16))
def test_gemm_case_fp32x3(cuda, idx):
    """fp32x3 mode (A_lo / B_lo / D_lo): full fp32 operands as (hi, lo) tf32 pairs, three tensor-core passes; the result
    (D + D_lo where the epilogue rounds) must match the fp64 product of the UNSPLIT operands to fp32-level accuracy."""
    import gemm_cases as G
    import torch
    from mtrl_b200 import _lib as L

    case = G.CASES[idx]
    if case[6] in (L.EPI_BIAS_RELU, L.EPI_RELU_MASK, L.EPI_STORE_TF32) and case[7] % 32 != 0:
        rel, D, _ = G.run_case_x3(case, with_d_lo=False)   # D_lo needs the TMA output path; D alone is tf32-rounded
        assert rel < 5e-4, f"{case[0]}: rel err {rel}"
        return
    rel, D, D_lo = G.run_case_x3(case)
    assert rel < 3e-6, f"{case[0]}: rel err {rel}"
    if D_lo is not None:   # D is exactly tf32 and D_lo its (tf32) remainder
        assert torch.equal(G.tf32_round(D), D) and torch.equal(G.tf32_round(D_lo), D_lo)
        assert float(D_lo.abs().max()) <= float(D.abs().max()) * 2.0 ** -11


def test_gemm_grouped_mixed_precision(cuda):
    """One launch mixing tf32 and fp32x3 problems (with and without D_lo): the staging-box bookkeeping of the epilogue."""
    import gemm_cases as G
    import torch
    from mtrl_b200 import _lib as L

    c = G.CASES
    p1, D1, ref1, k1 = G.make_problem(*c[3][1:], seed=21)
    p2, D2, D2lo, ref2, k2 = G.make_problem_x3(*c[3][1:], seed=22)
    p3, D3, ref3, k3 = G.make_problem(*c[8][1:], seed=23)
    p4, D4, D4lo, ref4, k4 = G.make_problem_x3(*c[8][1:], seed=24)
    p5, D5, _, ref5, k5 = G.make_problem_x3(*c[7][1:], seed=25)
    L.GemmPlan([p1, p2, p3, p4, p5]).run()
    torch.cuda.synchronize()
    r = lambda out, ref: float((out.double() - ref).norm() / ref.norm())  # noqa: E731
    assert r(D1, ref1) < 5e-4 and r(D3, ref3) < 5e-4
    assert r(D2.double() + D2lo.double(), ref2) < 3e-6 and r(D4.double() + D4lo.double(), ref4) < 3e-6
    assert r(D5, ref5) < 3e-6


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


@pytest.mark.parametrize("rows", [2304, 640])
@pytest.mark.parametrize("flags", [0, 32])
@pytest.mark.parametrize("ctas", [2, 1])
def test_phased_launch_matches_layer_by_layer(cuda, ctas, flags, rows):
    """Three dependent Dense layers (each reads what the previous one wrote) as ONE launch -- with grid-wide phase barriers
    (`problem.phase`), or with per-row-tile dependencies (MTRL_GEMM_ROWDEPS = 32: a tile of layer l + 1 starts once its rows of
    layer l are stored, no barrier) -- vs one launch per layer: bit-identical outputs, repeatedly (a race on the barrier / the
    counters / the TMA-store visibility would show up as a stale operand in some repetition).  640 rows: fewer tiles than SMs,
    so most workers run ahead into the next layer and really wait on the counters."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, W, K0 = rows, 1024, 96
    g = torch.Generator().manual_seed(7)
    X = G.tf32_round(torch.randn(M, K0, generator=g)).cuda()
    Ws = [G.tf32_round(torch.randn(K0 if i == 0 else W, W, generator=g) * (2.0 / (K0 if i == 0 else W)) ** 0.5).cuda() for i in range(3)]
    bs = [(torch.randn(W, generator=g) * 0.1).cuda() for _ in range(3)]

    def problems(outs, phased):
        ps, x = [], X
        for i in range(3):
            ps.append(L.GemmProblem(A=x.data_ptr(), lda=x.stride(0), a_major=0, B=Ws[i].data_ptr(), ldb=W, b_major=1,
                                    D=outs[i].data_ptr(), ldd=W, M=M, N=W, K=x.shape[1], block_n=256, k_splits=1,
                                    epilogue=L.EPI_BIAS_RELU, bias=bs[i].data_ptr(), phase=i if phased else 0))
            x = outs[i]
        return ps

    ref = [torch.zeros(M, W, device="cuda") for _ in range(3)]
    for p in problems(ref, False):
        L.GemmPlan([p], ctas=ctas).run()
    torch.cuda.synchronize()
    assert float(ref[2].abs().sum()) > 0
    out = [torch.zeros(M, W, device="cuda") for _ in range(3)]
    plan = L.GemmPlan(problems(out, True), ctas=ctas | flags)
    for rep in range(40):
        for o in out:
            o.fill_(float("nan"))
        plan.run()
        torch.cuda.synchronize()
        for i in range(3):
            assert torch.equal(out[i], ref[i]), f"repetition {rep}: layer {i} differs"


def test_phased_backward_chain(cuda):
    """dX of layer l feeds dW and dX of layer l - 1 inside one launch (MN-major 3-D maps read what a TMA store of the
    previous phase wrote)."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, W = 1536, 768
    g = torch.Generator().manual_seed(9)
    dZ2 = G.tf32_round(torch.randn(M, W, generator=g)).cuda()
    Hs = [G.tf32_round(torch.relu(torch.randn(M, W, generator=g))).cuda() for _ in range(2)]   # H0, H1 (forward activations)
    Wk = [G.tf32_round(torch.randn(W, W, generator=g) * (2.0 / W) ** 0.5).cuda() for _ in range(3)]  # kernels (in, out) of layers 1, 2

    def build(dZ1, dZ0, dW2, dW1, phased):
        def dx(src, Wl, mask, dst, ph):
            return L.GemmProblem(A=src.data_ptr(), lda=W, a_major=0, B=Wl.data_ptr(), ldb=W, b_major=0, D=dst.data_ptr(), ldd=W,
                                 M=M, N=W, K=W, block_n=256, k_splits=1, epilogue=L.EPI_RELU_MASK, mask=mask.data_ptr(), ldmask=W,
                                 phase=ph if phased else 0)

        def dw(Xl, dZl, dst, ph):
            return L.GemmProblem(A=Xl.data_ptr(), lda=W, a_major=1, B=dZl.data_ptr(), ldb=W, b_major=1, D=dst.data_ptr(), ldd=W,
                                 M=W, N=W, K=M, block_n=256, k_splits=1, epilogue=L.EPI_STORE, phase=ph if phased else 0)
        return [[dw(Hs[1], dZ2, dW2, 0), dx(dZ2, Wk[2], Hs[1], dZ1, 0)], [dw(Hs[0], dZ1, dW1, 1), dx(dZ1, Wk[1], Hs[0], dZ0, 1)]]

    z = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    r = [z(M, W), z(M, W), z(W, W), z(W, W)]
    for layer in build(*r, False):
        L.GemmPlan(layer).run()
    torch.cuda.synchronize()
    o = [z(M, W), z(M, W), z(W, W), z(W, W)]
    plan = L.GemmPlan(sum(build(*o, True), []))
    for rep in range(20):
        for t in o:
            t.fill_(float("nan"))
        plan.run()
        torch.cuda.synchronize()
        for a, b in zip(o, r):
            assert torch.equal(a, b), f"repetition {rep}"


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("hd", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(384, 512, 320, 256), (256, 400, 96, 224), (384, 256, 64, 128)])
def test_fused_output_head(cuda, ctas, hd, shape):
    """mtrl_gemm_problem_t::head_w: the own-task head (nn.vmap(Dense), multi_head.py:50-66) accumulated by the bias + ReLU
    epilogue, head_out[m][j] = sum_n D[m][n] head_w[task(m)][n][j], against a matmul of the D the launch stored."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, N, K, bn = shape
    p, D, ref, keep = G.make_problem(M, N, K, 0, 1, L.EPI_BIAS_RELU, bn, 1, seed=3)
    T = 3
    g = torch.Generator(device="cpu").manual_seed(7)
    head_w = torch.randn(T, N, hd, generator=g).cuda()
    tile_task = torch.tensor([(2 * i + 1) % T for i in range(M // 128)], dtype=torch.int32, device="cuda")
    out = torch.zeros(M, hd, device="cuda")
    p.head_w, p.head_out, p.head_tile_task, p.head_dim = head_w.data_ptr(), out.data_ptr(), tile_task.data_ptr(), hd
    L.GemmPlan([p], ctas=ctas).run()
    torch.cuda.synchronize()
    rel, _ = G.rel_err(D, ref, L.EPI_BIAS_RELU)
    assert rel < 5e-4
    want = torch.cat([D[128 * i:128 * (i + 1)].double() @ head_w[int(tile_task[i])].double() for i in range(M // 128)])
    err = float((out.double() - want).norm() / want.norm())
    assert err < 2e-6, f"fused head rel err {err}"


def test_fused_output_head_fp32x3(cuda):
    """With D_lo the head sees the unrounded result (hi + lo), as the stand-alone head kernels do in that mode."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, N, K, hd = 256, 512, 320, 8
    p, D, D_lo, ref, keep = G.make_problem_x3(M, N, K, 0, 1, L.EPI_BIAS_RELU, 128, 1, seed=5)
    head_w = torch.randn(2, N, hd, generator=torch.Generator().manual_seed(9)).cuda()
    tile_task = torch.tensor([1, 0], dtype=torch.int32, device="cuda")
    out = torch.zeros(M, hd, device="cuda")
    p.head_w, p.head_out, p.head_tile_task, p.head_dim = head_w.data_ptr(), out.data_ptr(), tile_task.data_ptr(), hd
    L.GemmPlan([p]).run()
    torch.cuda.synchronize()
    full = D.double() + D_lo.double()
    want = torch.cat([full[128 * i:128 * (i + 1)] @ head_w[int(tile_task[i])].double() for i in range(2)])
    err = float((out.double() - want).norm() / want.norm())
    assert err < 2e-6, f"fused head (fp32x3) rel err {err}"


# ---------------------------------------------------------------------------------------------
# Stream-K plans (MTRL_GEMM_STREAMK): k-ranges instead of whole tiles; cut tiles are finished by the last unit to arrive.
# ---------------------------------------------------------------------------------------------
STREAMK_EXTRA = [
    # few tiles, long K: every tile is cut into many pieces (store, bias + ReLU, ReLU mask, rounding store)
    ("sk_kk_store_longk", 256, 256, 4096, 0, 0, 0, 256, 1),
    ("sk_fwd_longk", 384, 512, 2048, 0, 1, 1, 256, 1),
    ("sk_dx_mask_longk", 384, 512, 2048, 0, 0, 2, 256, 1),
    ("sk_store_tf32", 256, 384, 1024, 0, 0, 4, 128, 1),
    # a shard-like launch: 7 row tiles x 8 column tiles, K = 2048
    ("sk_shard_fwd", 896, 2048, 2048, 0, 1, 1, 256, 1),
    ("sk_dw_splitk", 512, 512, 1280, 1, 1, 3, 256, 3),
]


@pytest.mark.parametrize("ctas", [2, 1])
@pytest.mark.parametrize("idx", range(16 + len(STREAMK_EXTRA)))
def test_gemm_case_streamk(cuda, idx, ctas):
    import gemm_cases as G
    from mtrl_b200 import _lib as L

    case = (G.CASES + STREAMK_EXTRA)[idx]
    for rep in range(2):   # the second launch runs on the counters the first one re-armed
        rel, _ = G.run_case(case, ctas=ctas | L.GEMM_STREAMK)
        assert rel < 5e-4, f"{case[0]} (stream-K, ctas {ctas}, launch {rep}): rel err {rel}"


@pytest.mark.parametrize("ctas", [2, 1])
def test_streamk_repeated_launches_and_fused_extras(cuda, ctas):
    """One stream-K plan launched repeatedly (counters re-armed by the finishers), with everything the fused epilogues emit:
    ReLU bits, the fused output head (accumulated once per tile, by the finisher only), column-sum partials."""
    import torch

    import gemm_cases as G
    from mtrl_b200 import _lib as L

    M, N, K, hd = 640, 512, 1536, 8
    p, D, ref, keep = G.make_problem(M, N, K, 0, 1, L.EPI_BIAS_RELU, 256, 1, seed=11)
    head_w = torch.randn(3, N, hd, generator=torch.Generator().manual_seed(2)).cuda()
    tile_task = torch.tensor([i % 3 for i in range(M // 128)], dtype=torch.int32, device="cuda")
    out = torch.zeros(M, hd, device="cuda")
    bits = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda")
    p.head_w, p.head_out, p.head_tile_task, p.head_dim = head_w.data_ptr(), out.data_ptr(), tile_task.data_ptr(), hd
    p.relu_bits_out, p.ldbits = bits.data_ptr(), N // 32
    q, Dq, refq, keepq = G.make_problem(M, N, K, 0, 0, L.EPI_RELU_MASK, 256, 1, seed=12)
    cs = torch.zeros((M + 31) // 32, N, device="cuda")
    q.colsum_partial = cs.data_ptr()
    plan = L.GemmPlan([p, q], ctas=ctas | L.GEMM_STREAMK)
    for _ in range(3):   # (which unit finishes a tile depends on arrival order, so launches agree to rounding, not bit for bit)
        out.zero_()
        plan.run()
        torch.cuda.synchronize()
        assert G.rel_err(D, ref, L.EPI_BIAS_RELU)[0] < 5e-4 and G.rel_err(Dq, refq, L.EPI_RELU_MASK)[0] < 5e-4
        want = torch.cat([D[128 * i:128 * (i + 1)].double() @ head_w[int(tile_task[i])].double() for i in range(M // 128)])
        assert float((out.double() - want).norm() / want.norm()) < 2e-6
        want_bits = (D > 0).view(M, N // 32, 32).to(torch.int64).mul(2 ** torch.arange(32, device="cuda")).sum(-1)
        assert torch.equal(bits.to(torch.int64) & 0xFFFFFFFF, want_bits)
        want_cs = Dq.view((M + 31) // 32, 32, N).double().sum(1)
        assert float((cs.double() - want_cs).norm() / want_cs.norm()) < 1e-5


def test_streamk_fp32x3(cuda):
    import gemm_cases as G
    from mtrl_b200 import _lib as L

    for case in (G.CASES[3], G.CASES[8], ("sk_x3_longk", 256, 384, 2048, 0, 1, 1, 128, 1)):
        rel, _, _ = G.run_case_x3(case, ctas=2 | L.GEMM_STREAMK)
        assert rel < 3e-6, f"{case[0]} fp32x3 stream-K: rel err {rel}"


def test_rowdeps_rejects_what_needs_a_barrier(cuda):
    """A row-dependency plan takes chains of Dense layers only: reading an earlier output as the B operand (dW of a backward
    chain) needs the whole previous problem and is refused at plan creation."""
    import torch

    from mtrl_b200 import _lib as L

    M, W = 256, 256
    a, w, d0, d1 = (torch.zeros(M, W, device="cuda") for _ in range(4))
    p0 = L.GemmProblem(A=a.data_ptr(), lda=W, a_major=0, B=w.data_ptr(), ldb=W, b_major=0, D=d0.data_ptr(), ldd=W, M=M, N=W, K=W,
                       block_n=256, k_splits=1, epilogue=L.EPI_STORE, phase=0)
    p1 = L.GemmProblem(A=a.data_ptr(), lda=W, a_major=1, B=d0.data_ptr(), ldb=W, b_major=1, D=d1.data_ptr(), ldd=W, M=W, N=W, K=M,
                       block_n=256, k_splits=1, epilogue=L.EPI_STORE, phase=1)
    with pytest.raises(L.MtrlError):
        L.GemmPlan([p0, p1], ctas=2 | L.GEMM_ROWDEPS)
