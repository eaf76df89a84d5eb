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
