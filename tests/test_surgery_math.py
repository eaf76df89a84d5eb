"""CPU: the coefficient-space formulations the CUDA kernels use for PCGrad / CAGrad / GradNorm (csrc/sac_kernels.cuh
pcgrad_coeff_kernel, cagrad_coeff_kernel, gradnorm_coeff_kernel -- everything from the T x T Gram matrix, then one
weighted sum of the rows) against the literal restatements of mtrl/optim/{pcgrad,cagrad,gradnorm}.py in
oracle/taskgrad_oracle.py, on random per-task gradient matrices.  The ports below follow the kernels line by line."""
import math

import pytest
import torch

from oracle import taskgrad_oracle as TG


def rows_and_gram(T, P, seed, conflict=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, P, generator=g, dtype=torch.float64) * (10.0 ** torch.linspace(-1, 1.5, T, dtype=torch.float64))[:, None]
    if conflict and T >= 3:
        x[1] = -0.3 * x[0] + 0.05 * torch.randn(P, generator=g, dtype=torch.float64)
        x[2] = -0.5 * x[1] + 0.2 * x[2]
    rows = x / T                      # what the kernels see: full-batch-mean gradients restricted to one task's rows
    return x, rows, rows @ rows.T


def pcgrad_weights(gram, T, perm):
    """pcgrad_coeff_kernel: projections in coefficient space over the permuted, rescaled Gram matrix."""
    G = gram[perm][:, perm] * float(T * T)
    C = torch.eye(T, dtype=torch.float64)
    conflicts = 0
    for i in range(T):
        for j in range(T):
            proj = (C[i] @ G[:, j]) / (G[j, j] + 1e-8)
            if proj < 0:
                C[i, j] -= proj
                conflicts += 1
    w = torch.zeros(T, dtype=torch.float64)
    w[perm] = C.sum(dim=0)            # mean over tasks of T x (unscaled rows)
    after = torch.sqrt(torch.clamp(torch.einsum("ia,ab,ib->i", C, G, C), min=0)).mean()
    return w, conflicts / 2, after


@pytest.mark.parametrize("T,P,seed", [(3, 50, 0), (6, 400, 1), (10, 257, 2), (50, 300, 3)])
def test_pcgrad_in_coefficient_space(T, P, seed):
    x, rows, gram = rows_and_gram(T, P, seed)
    perm = torch.randperm(T, generator=torch.Generator().manual_seed(seed + 100))
    ref, stats = TG.pcgrad(x, perm)
    w, conflicts, after = pcgrad_weights(gram, T, perm)
    got = (w[:, None] * rows).sum(dim=0)
    assert conflicts == stats["n_grad_conflicts"] and conflicts > 0
    assert torch.allclose(got, ref, rtol=1e-9, atol=1e-12)
    assert abs(float(after) - float(stats["avg_grad_magnitude"])) <= 1e-9 * float(stats["avg_grad_magnitude"])


def cagrad_weights(gram, T, c=0.5, iters=21, momentum=0.5):
    """cagrad_coeff_kernel (double precision, analytic gradient of the objective)."""
    gs = float(T * T)
    lr = 25.0 if T < 50 else 50.0
    n = torch.sqrt(torch.clamp(torch.diagonal(gram) * gs, min=0))
    clipc = torch.clamp(1.0 / (n + 1e-8), max=1.0)
    GG = gram * gs * clipc[:, None] * clipc[None, :]
    scale = torch.sqrt(torch.diagonal(GG) + 1e-4).mean()
    GG = GG / scale ** 2
    Gg = GG.mean(dim=1)
    cn = math.sqrt(float(Gg.mean()) + 1e-4) * c

    def objective(w, want_grad):
        s = 1e-8 + w.sum()
        ww = w / s
        d = GG @ ww
        root = math.sqrt(float(ww @ d) + 1e-4)
        val = float(ww @ Gg) + cn * root
        if not want_grad:
            return val, None
        d = Gg + cn * d / root
        return val, (d - (d @ ww)) / s

    w = torch.zeros(T, dtype=torch.float64)
    vel = torch.zeros(T, dtype=torch.float64)
    wb, ob = w.clone(), float("inf")
    for _ in range(iters - 1):
        o, g = objective(w, True)
        if o < ob:
            ob, wb = o, w.clone()
        vel = momentum * vel + g
        w = w - lr * vel
    o, _ = objective(w, False)
    if o < ob:
        ob, wb = o, w.clone()
    tw = torch.softmax(wb, dim=0)
    lmbda = cn / (math.sqrt(float(tw @ GG @ tw) + 1e-4) + 1e-4)
    return (1.0 / T + tw * lmbda) / (1 + c * c) * clipc * math.sqrt(gs), tw, ob


@pytest.mark.parametrize("T,P,seed", [(3, 50, 0), (6, 400, 1), (50, 300, 3)])
def test_cagrad_from_the_gram_matrix(T, P, seed):
    x, rows, gram = rows_and_gram(T, P, seed)
    ref, stats = TG.cagrad(x)
    w, tw, ob = cagrad_weights(gram, T)
    got = (w[:, None] * rows).sum(dim=0)
    assert torch.allclose(tw, stats["task_weights"], atol=1e-9)
    assert abs(ob - float(stats["cagrad_objective"])) <= 1e-9
    assert torch.allclose(got, ref, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("clip", [False, True])
def test_gradnorm_reduces_to_the_sum_of_clipped_gradients(clip):
    T, P = 5, 120
    x, rows, gram = rows_and_gram(T, P, 7)
    ref, stats = TG.gradnorm(x, torch.rand(T, dtype=torch.float64) + 0.5, max_grad_norm=1.0 if clip else None)
    assert torch.allclose(stats["task_weights"], torch.ones(T, dtype=torch.float64))
    n = torch.sqrt(torch.diagonal(gram) * T * T)
    w = (torch.clamp(1.0 / (n + 1e-8), max=1.0) if clip else torch.ones(T, dtype=torch.float64)) * T   # gradnorm_coeff_kernel
    assert torch.allclose((w[:, None] * rows).sum(dim=0), ref, rtol=1e-9, atol=1e-12)


def test_support_and_elementwise_metrics_agree_with_brute_force():
    """oracle/taskgrad_oracle.py's vectorised restatements of compute_sparsity_mismatch / compute_support_metrics
    against element-by-element loops (utils.py:75-91, mtsac.py:804-835)."""
    x, _, _ = rows_and_gram(4, 60, 11)
    x = x / 5
    em, sm = TG.elementwise_metrics(x), TG.support_metrics(x)
    thr = [torch.quantile(x[t].abs(), 0.8) for t in range(4)]
    for a in range(4):
        nz = (x[a].abs() < 1e-3)
        for b in range(4):
            mism = sum(1 for p in range(60) if nz[p] and abs(x[b, p]) > 1.0)
            expect = 0.0 if a == b else mism / max(int(nz.sum()), 1)
            assert abs(float(em["pairwise_interference_rate"][a, b]) - expect) < 1e-12
            sa, sb = x[a].abs() >= thr[a], x[b].abs() >= thr[b]
            inter, union = int((sa & sb).sum()), int((sa | sb).sum())
            assert abs(float(sm["pairwise_jaccard"][a, b]) - inter / (union + 1e-8)) < 1e-9
            conf = (x[a] * x[b]) < 0
            genuine, ghost = int((sa & sb & conf).sum()), int((~(sa & sb) & conf).sum())
            assert abs(float(sm["genuine_conflict_rate"][a, b]) - genuine / (genuine + ghost + 1e-8)) < 1e-9
