"""Full-size checks at BASELINE.json's configs[2] shape (synthetic ml-20m: 138,493 x 26,744, 20 M interactions,
K = 200, B = 4096) through size-independent properties: the tcgen05 3xTF32 path against the exact-fp32 FFMA path on
the same inputs, padding invariants, (score desc, iid desc) order / novelty / idempotence of the full-catalog top-k,
sampler membership properties.  (Parity against the oracle at this size: tests/test_gpu_baseline_shapes.py.)"""
import numpy as np
import pytest

import drecpy_b200 as drb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c3():
    u, i, v = drb.synthetic_interactions(138493, 26744, 20_000_000, seed=10, zipf_a=1.0)
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    return ds


def test_full_size_step_tcgen05_equals_ffma_and_invariants(c3):
    import torch
    ds, B = c3, 4096
    models = []
    for gemm in ('tcgen05', 'ffma'):
        m = drb.CDAE(hidden_factors=200, corruption_level=0.2, seed=10, verbose=False, rng_mode='philox', gemm=gemm)
        m.fit(ds, epochs=0, batch_size=B, learning_rate=1e-3, reg_rate=1e-3)
        models.append(m)
    assert ds.count_unique('uid') == 138493 and ds.count_unique('iid') == 26744
    sampler = drb.PointSampler(ds, 5, 1e-3, 10)
    deg = np.diff(ds.csr(1e-3)[0])
    prev = None
    for step in range(3):
        uids, iids, vals = sampler.sample_arrays(B)
        # sampler properties: null pairs are absent from the data, positive pairs are present with their value
        neg = vals == 0
        indptr, indices, data = ds.csr()
        for r in np.flatnonzero(neg)[:50]:
            assert iids[r] not in indices[indptr[uids[r]]:indptr[uids[r] + 1]]
        for r in np.flatnonzero(~neg)[:50]:
            row = indices[indptr[uids[r]]:indptr[uids[r] + 1]]
            assert data[indptr[uids[r]] + np.searchsorted(row, iids[r])] == vals[r]
        off = np.concatenate([[0], np.cumsum(deg[uids])]).astype(np.int32)
        d_u, d_o = torch.as_tensor(uids, device='cuda'), torch.as_tensor(off, device='cuda')
        losses = []
        for m in models:
            loss = torch.zeros(2, device='cuda')
            m.step_device(d_u, d_o, None, 1e-3, loss)
            losses.append(loss.cpu().numpy())
        assert abs(losses[0][0] - losses[1][0]) <= 2e-5 * abs(losses[1][0]), (step, losses)
        assert np.isfinite(losses[0]).all() and losses[0][1] > 0
        if prev is not None:
            assert losses[0][0] < prev                       # the loss goes down on this data
        prev = losses[0][0]
    a, b = models[0]._params, models[1]._params
    assert float((a - b).abs().max()) <= 5e-4 * float(b.abs().max())
    L = models[0]._L
    for m in models:                                         # padding of the item-major tables stays exactly zero
        w2t = m._params[L.off_w2t:L.off_w2t + 26744 * L.ld].view(26744, L.ld)
        assert L.ld == 200 or float(w2t[:, 200:].abs().max()) == 0.0
        assert float(m._params[L.off_b2 + 26744:L.off_v].abs().sum()) == 0.0


def test_full_size_topk_properties(c3):
    import torch
    ds = c3
    m = drb.CDAE(hidden_factors=200, seed=10, verbose=False, rng_mode='philox')
    m.fit(ds, epochs=2, batch_size=4096)
    uids = np.arange(0, 138493, 97, dtype=np.int32)
    oi, os_, on = m.topk_batch(uids, 100, novelty=True)
    oi2, os2, on2 = m.topk_batch(uids, 100, novelty=True)
    assert np.array_equal(oi, oi2) and np.array_equal(os_, os2) and np.array_equal(on, on2)     # idempotent
    assert (on == 100).all()
    # (score desc, iid desc) order, i.e. heapq.nlargest over (score, iid) tuples (cdae.py:102-103)
    assert (os_[:, :-1] >= os_[:, 1:]).all()
    ties = os_[:, :-1] == os_[:, 1:]
    assert (oi[:, :-1][ties] > oi[:, 1:][ties]).all()
    indptr, indices, _ = ds.csr()
    for r in range(0, len(uids), 53):                       # novelty: nothing the user has already interacted with
        seen = indices[indptr[uids[r]]:indptr[uids[r] + 1]]
        assert len(np.intersect1d(seen, oi[r])) == 0
        p = m._predict(int(uids[r]))                         # scores are the dense predictions, and they are the top ones
        assert np.allclose(p[oi[r]], os_[r], rtol=1e-5)      # tensor-core scores vs the exact-fp32 predict path
        p[seen] = -1
        assert np.sort(p)[-100] <= os_[r, -1] * (1 + 1e-5)
