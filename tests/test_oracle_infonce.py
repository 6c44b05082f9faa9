"""Known-answer tests that pin the AUTHORED symmetric-InfoNCE oracle (no reference implementation
exists -- SURVEY.md section 0 / section 8c)."""
import math

import torch

from oracle import infonce as oi


def test_orthonormal_rows_closed_form():
    B, D, tau = 64, 128, 0.07
    q, _ = torch.linalg.qr(torch.randn(D, D, dtype=torch.float64))
    e = q[:B] * 3.0  # any positive row scale is removed by the normalisation
    loss = oi.symmetric_infonce(e, e.clone(), tau)
    expect = math.log(1 + (B - 1) * math.exp(-1 / tau))
    assert abs(float(loss) - expect) < 1e-12


def test_all_equal_embeddings_give_log_B():
    B = 37
    e = torch.ones(B, 16, dtype=torch.float64)
    assert abs(float(oi.symmetric_infonce(e, e, 0.07)) - math.log(B)) < 1e-12


def test_similarity_matrix_is_scaled_cosine():
    e, f = torch.randn(5, 8, dtype=torch.float64), torch.randn(7, 8, dtype=torch.float64)
    S = oi.similarity_matrix(e, f, 0.5)
    assert S.shape == (5, 7)
    cos = torch.nn.functional.cosine_similarity(e[:, None, :], f[None, :, :], dim=-1)
    assert torch.allclose(S, cos / 0.5, atol=1e-12)
    assert float(S.abs().max()) <= 2.0 + 1e-12


def test_closed_form_gradient_and_gradcheck():
    torch.manual_seed(0)
    e = torch.randn(6, 5, dtype=torch.float64, requires_grad=True)
    f = torch.randn(6, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: oi.symmetric_infonce(a, b, 0.3), (e, f), eps=1e-6, atol=1e-6)
    S = oi.similarity_matrix(e, f, 0.3).detach().requires_grad_(True)
    t = torch.arange(6)
    L = 0.5 * (torch.nn.functional.cross_entropy(S, t) + torch.nn.functional.cross_entropy(S.t(), t))
    (g,) = torch.autograd.grad(L, S)
    assert torch.allclose(g, oi.infonce_grad_S(S.detach()), atol=1e-12)


def test_sharded_loss_parts_sum_to_global():
    torch.manual_seed(1)
    e, f = torch.randn(12, 8, dtype=torch.float64), torch.randn(12, 8, dtype=torch.float64)
    total, parts = oi.sharded_symmetric_infonce(e.chunk(3), f.chunk(3), 0.07)
    assert abs(float(total) - float(oi.symmetric_infonce(e, f, 0.07))) < 1e-12
    assert abs(float(sum(parts)) - float(total)) < 1e-12
