"""GPU: est_quad_linear_robust drop-in (csrc/irls.cu) against the reference's own output (golden irls_400.npz, written by
oracle/pin_against_reference.py from util/transform_estimation.py) and against the oracle restatement with weights."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(T, T_ref):
    T, T_ref = torch.as_tensor(T).double().cpu(), torch.as_tensor(T_ref).double()
    assert float(torch.linalg.norm(T[:3, :3] - T_ref[:3, :3])) < 1e-4
    assert float(torch.linalg.norm(T[:3, 3] - T_ref[:3, 3])) < 1e-3
    assert torch.equal(T[3], torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64))


def test_golden_reference_output(golden_dir):
    from eyoc_b200.util.transform_estimation import est_quad_linear_robust
    g = np.load(f'{golden_dir}/irls_400.npz')
    p0, p1 = torch.from_numpy(g['p0']), torch.from_numpy(g['p1'])
    T = est_quad_linear_robust(p0, p1)                       # CPU tensors in, like the reference's callers
    assert T.device.type == 'cpu' and T.shape == (4, 4)
    _close(T, g['T'])
    _close(est_quad_linear_robust(p0.cuda(), p1.cuda()), g['T'])


def test_weighted_with_outliers_vs_oracle():
    from eyoc_b200.util.transform_estimation import est_quad_linear_robust
    from oracle import matching_oracle as MO
    g = torch.Generator().manual_seed(3)
    n = 5000
    p0 = torch.randn(n, 3, generator=g) * torch.tensor([30.0, 30.0, 2.0])
    ang = torch.tensor(0.05)
    R = torch.tensor([[torch.cos(ang), -torch.sin(ang), 0.0], [torch.sin(ang), torch.cos(ang), 0.0], [0.0, 0.0, 1.0]])
    p1 = p0 @ R.T + torch.tensor([0.4, -0.2, 0.05]) + torch.randn(n, 3, generator=g) * 0.02
    p1[::10] += torch.randn(n // 10, 3, generator=g) * 5.0       # 10 % outliers
    w = torch.rand(n, 1, generator=g) * 0.5 + 0.5
    _close(est_quad_linear_robust(p0.cuda(), p1.cuda(), w.cuda()), MO.irls_pose(p0, p1, w))
    _close(est_quad_linear_robust(p0, p1), MO.irls_pose(p0, p1))
