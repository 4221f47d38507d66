"""GPU: tensor-core pre-filtered 1-NN (eyoc_knn1_tc) == the fp32-FMA kernel (eyoc_knn1), bit for bit - indices and
values - on unit descriptors, unnormalised data with a wide range of norms, clustered near-duplicates, exact duplicates
(lowest index wins), ragged sizes, and batches that must fall back (NaN / Inf / fp16 overflow)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _both(F0, F1, form):
    from eyoc_b200.lib import eval as ev
    saved = ev.KNN_MODE
    try:
        ev.KNN_MODE = 'fp32'
        i0, d0 = ev.knn1(F0, F1, form=form, return_distance=True)
        ev.KNN_MODE = 'tc'
        i1, d1 = ev.knn1(F0, F1, form=form, return_distance=True)
    finally:
        ev.KNN_MODE = saved
    torch.cuda.synchronize()
    return i0, d0, i1, d1


def _check(F0, F1, form):
    i0, d0, i1, d1 = _both(F0, F1, form)
    assert torch.equal(i0, i1), f'{int((i0 != i1).sum())} of {i0.numel()} indices differ'
    assert torch.equal(d0.view(torch.int32), d1.view(torch.int32))      # bit-equal, NaN included


def _unit(g, *shape):
    x = torch.randn(*shape, generator=g)
    return (x / x.norm(dim=-1, keepdim=True)).cuda()


@pytest.mark.parametrize('form', [0, 1])
@pytest.mark.parametrize('B,nq,nr', [(1, 5000, 5000), (2, 8000, 8000), (3, 1000, 777), (1, 1, 1), (2, 129, 257), (1, 300, 5)])
def test_unit_descriptors(form, B, nq, nr):
    g = torch.Generator().manual_seed(B * 100000 + nq + nr + form)
    _check(_unit(g, B, nq, 32), _unit(g, B, nr, 32), form)


@pytest.mark.parametrize('form', [0, 1])
def test_wide_norm_range_and_offsets(form):
    g = torch.Generator().manual_seed(7 + form)
    q = torch.randn(2, 3000, 32, generator=g) * torch.logspace(-3, 2, 3000)[None, :, None]
    r = torch.randn(2, 4100, 32, generator=g) * torch.logspace(-2, 1.5, 4100)[None, :, None] + 0.3
    _check(q.cuda(), r.cuda(), form)


@pytest.mark.parametrize('scale', [1.05, 1.5, 2.0])
def test_form1_products_above_one_give_first_nan(scale):
    """Un-normalised descriptors with norms just above 1 .. around 2: wherever <q, r> > 1 the compared value
    sqrt(2 - 2 s + 1e-6) is NaN and torch.argmin / the fp32 kernel return the FIRST such column whatever its s - the
    pre-filter must not drop an earlier NaN column for lying far below the row maximum."""
    g = torch.Generator().manual_seed(int(scale * 100))
    q = _unit(g, 2, 1500, 32) * scale
    r = _unit(g, 2, 2300, 32) * scale
    # a few strongly aligned references per query at scattered indices, so several columns exceed 1 by different margins
    for b in range(2):
        for k in range(0, 1500, 3):
            j = (k * 7 + 11) % 2300
            r[b, j] = q[b, k] * (0.7 + 0.3 * ((k % 5) / 4.0))
    i0, d0, i1, d1 = _both(q, r, 1)
    assert bool(torch.isnan(d0).any())                              # the case is exercised
    assert torch.equal(i0, i1), f'{int((i0 != i1).sum())} of {i0.numel()} indices differ'
    assert torch.equal(d0.view(torch.int32), d1.view(torch.int32))


@pytest.mark.parametrize('form', [0, 1])
def test_clustered_near_duplicates_and_exact_ties(form):
    """Many reference rows within 1e-4 .. 1e-7 of each other (the pre-filter must keep every possible winner), plus exact
    duplicates at different indices (the lowest index must win)."""
    g = torch.Generator().manual_seed(11 + form)
    base = _unit(g, 1, 200, 32).cpu()
    reps = base.repeat(1, 30, 1)                                   # 6000 rows: 30 copies of 200 centres
    noise = torch.randn(1, 6000, 32, generator=g) * torch.logspace(-7, -3, 6000)[None, :, None]
    noise[:, ::7] = 0.0                                            # exact duplicates
    r = reps + noise
    r = r / r.norm(dim=-1, keepdim=True)
    r[:, ::7] = reps[:, ::7]                                       # keep those bit-identical
    q = base.repeat(1, 10, 1) + torch.randn(1, 2000, 32, generator=g) * 1e-5
    q = q / q.norm(dim=-1, keepdim=True)
    _check(q.cuda(), r.cuda(), form)
    i0, _, i1, _ = _both(r[:, :512].cuda(), r.cuda(), form)        # queries that ARE reference rows
    assert torch.equal(i0, i1)


@pytest.mark.parametrize('form', [0, 1])
def test_fallback_batches(form):
    """A batch holding NaN / Inf / values beyond the fp16 range is routed to the fp32-FMA kernel on the device; the other
    batches of the same call still take the tensor-core path."""
    g = torch.Generator().manual_seed(23 + form)
    q, r = _unit(g, 4, 700, 32), _unit(g, 4, 900, 32)
    q[1, 5, 3] = float('nan')
    r[2, 17, 0] = float('inf')
    r[3, 100] *= 1.0e5
    _check(q, r, form)


def test_default_mode_is_tc_and_matches_golden():
    """The drop-in symbols take the tensor-core path by default and reproduce the reference's own outputs (golden fixture
    written by oracle/pin_against_reference.py from lib/eval.py find_nn_gpu and SC2_PCR.py match_pair arithmetic)."""
    import os
    from eyoc_b200.lib import eval as ev
    assert ev.KNN_MODE == 'tc'
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'knn_1500x1300.npz'))
    F0, F1 = torch.from_numpy(gold['F0']).cuda(), torch.from_numpy(gold['F1']).cuda()
    idx = ev.knn1(F0, F1, form=0)
    assert np.array_equal(idx.cpu().numpy(), gold['idx_sq'])
    idx = ev.knn1(F0, F1, form=1)
    assert np.array_equal(idx.cpu().numpy(), gold['idx_cos'])
