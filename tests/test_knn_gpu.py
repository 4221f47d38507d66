"""GPU parity: fused kNN kernel (csrc/knn.cu) through the C-ABI vs the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _norm(x):
    return torch.nn.functional.normalize(x, dim=1)


def _classify(F0, F1, got, want, form):
    """Every index disagreement with the torch-order oracle must be an fp32 near-tie in fp64."""
    bad = np.nonzero(got != want)[0]
    a = F0[bad].astype(np.float64)
    g, w = F1[got[bad]].astype(np.float64), F1[want[bad]].astype(np.float64)
    if form == 0:
        dg, dw = ((a - g) ** 2).sum(1), ((a - w) ** 2).sum(1)
    else:
        dg, dw = -(a * g).sum(1), -(a * w).sum(1)
    return bad, np.abs(dg - dw)


def test_golden_knn(golden_dir):
    from eyoc_b200.lib.eval import find_nn_gpu, knn1
    g = np.load(f'{golden_dir}/knn_1500x1300.npz')
    F0, F1 = torch.from_numpy(g['F0']).cuda(), torch.from_numpy(g['F1']).cuda()
    inds, dists = find_nn_gpu(F0, F1, nn_max_n=500, return_distance=True)
    assert inds.device.type == 'cpu' and inds.dtype == torch.int64 and dists.shape == (1500, 1)
    bad, margin = _classify(g['F0'], g['F1'], inds.numpy(), g['idx_sq'], 0)
    assert len(bad) == 0 or margin.max() < 1e-6, (len(bad), margin)
    np.testing.assert_allclose(dists.numpy(), g['dist_sq'], rtol=0, atol=2e-6)
    idx_cos = knn1(F0, F1, form=1).cpu().numpy()
    bad, margin = _classify(g['F0'], g['F1'], idx_cos, g['idx_cos'], 1)
    assert len(bad) == 0 or margin.max() < 1e-6, (len(bad), margin)
    # exact duplicates planted at F1[100:140] == F1[200:240]: ties must resolve to the lower index
    assert not np.isin(inds.numpy(), np.arange(200, 240)).any()


@pytest.mark.parametrize('nq,nr,dim', [(1, 1, 32), (5, 3, 32), (129, 257, 32), (5000, 5000, 32), (8000, 8000, 32),
                                       (300, 200, 16), (300, 200, 7), (200, 300, 64), (77, 1000, 40)])
@pytest.mark.parametrize('form', [0, 1])
def test_kernel_order_bit_exact(nq, nr, dim, form):
    """Bit-exact (indices AND values) against the kernel-order oracle (sequential fp32 FMA)."""
    from eyoc_b200.lib.eval import knn1
    from oracle import matching_oracle as MO
    g = torch.Generator().manual_seed(nq * 31 + nr + dim)
    F0, F1 = _norm(torch.randn(nq, dim, generator=g)), _norm(torch.randn(nr, dim, generator=g))
    if nr > 40:
        F1[nr // 2: nr // 2 + 10] = F1[5:15]           # exact duplicates
    idx, dist = knn1(F0.cuda(), F1.cuda(), form=form, return_distance=True)
    want_i, want_d = (MO.knn_sq_seq if form == 0 else MO.knn_cos_seq)(F0.numpy(), F1.numpy())
    np.testing.assert_array_equal(idx.cpu().numpy(), want_i)
    np.testing.assert_array_equal(dist.cpu().numpy(), want_d)


def test_full_size_vs_torch_order():
    """BASELINE sizes: 5000x5000 (find_corr) and 8000x8000 (match_pair) against the torch-order oracle."""
    from eyoc_b200.lib.eval import knn1
    from oracle import matching_oracle as MO
    g = torch.Generator().manual_seed(5)
    for n, form in ((5000, 0), (8000, 1)):
        F0, F1 = _norm(torch.randn(n, 32, generator=g)), _norm(torch.randn(n, 32, generator=g))
        want = (MO.find_nn(F0, F1, nn_max_n=500) if form == 0 else MO.match_argmin(F0, F1)).numpy()
        got = knn1(F0.cuda(), F1.cuda(), form=form).cpu().numpy()
        bad, margin = _classify(F0.numpy(), F1.numpy(), got, want, form)
        assert len(bad) == 0 or margin.max() < 1e-6, (len(bad), margin)


def test_batched_equals_loop():
    from eyoc_b200.lib.eval import knn1
    g = torch.Generator().manual_seed(11)
    F0, F1 = _norm(torch.randn(3, 700, 32, generator=g)).cuda(), _norm(torch.randn(3, 900, 32, generator=g)).cuda()
    got = knn1(F0, F1, form=1)
    for b in range(3):
        assert torch.equal(got[b], knn1(F0[b], F1[b], form=1))


def test_nan_and_errors():
    from eyoc_b200.lib.eval import knn1, find_nn_gpu
    F0 = torch.randn(10, 32).cuda()
    F1 = torch.randn(20, 32).cuda()
    F1[7, 3] = float('nan')
    F1[12, 0] = float('nan')
    assert (knn1(F0, F1, form=0).cpu() == 7).all()      # torch.argmin: first NaN wins
    with pytest.raises(RuntimeError):
        knn1(F0, F1[:0], form=0)
    with pytest.raises(RuntimeError):
        find_nn_gpu(F0.cpu(), F1.cpu())                  # no CPU fallback
    assert knn1(F0[:0], F1, form=0).shape == (0,)
