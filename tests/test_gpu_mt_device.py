"""rng_mode='mt19937_device': the reference's corruption stream (cdae.py:63-64, one random.Random, n_items draws per
sampled user) replayed on the GPU by MT19937 jump-ahead -- bit-exact with the host replay and with CPython."""
import ctypes as C
import random

import numpy as np
import pytest

import drecpy_b200 as drb
from drecpy_b200 import _lib
from oracle.cdae import CDAEOracle, corruption_keep_mt
from oracle.sampler import PointSamplerOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('U,I,nnz,B,q', [(60, 97, 900, 6, 0.2), (943, 1682, 100000, 64, 0.2), (400, 5000, 30000, 300, 0.35),
                                         (300, 26744, 40000, 130, 0.2)])
def test_device_mask_stream_matches_host_replay(U, I, nnz, B, q):
    import torch
    from drecpy_b200.mask_stream import DeviceMaskStream
    u, i, v = drb.synthetic_interactions(U, I, nnz, seed=3)
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    indptr, indices, _ = ds.csr(0.001)
    indptr, indices = np.ascontiguousarray(indptr), np.ascontiguousarray(indices)
    lib = _lib.load()
    ctx = _lib.vp()
    _lib.check(lib.drb_ctx_create(0, C.byref(ctx)))
    _lib.check(lib.drb_ctx_set_stream(ctx, _lib.vp(torch.cuda.current_stream().cuda_stream)))
    host = _lib.HostRng(10)
    for _ in range(777):                               # start somewhere inside a 624-word block
        host.random()
    twin = _lib.HostRng(10)
    for _ in range(777):
        twin.random()
    d_indptr, d_indices = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    stream = DeviceMaskStream(twin, I, q, torch, torch.device('cuda'), ctx, d_indptr, d_indices)
    rng = np.random.default_rng(0)
    deg = np.diff(indptr)
    for step in range(4):
        uids = rng.integers(0, U, B).astype(np.int32)
        if step == 1:
            uids[: B // 2] = uids[0]                    # a repeated user: the stream still advances n_items draws per row
        off = np.zeros(B + 1, np.int32)
        keep = np.zeros(max(int(deg[uids].sum()), 1), np.uint8)
        _lib.check(lib.drb_cdae_corruption_keep_mt(host.handle, _lib.np_ptr(uids), B, I, float(q), _lib.np_ptr(indptr),
                                                   _lib.np_ptr(indices), _lib.np_ptr(off), _lib.np_ptr(keep), len(keep)))
        d_keep = torch.full((len(keep) + 16,), 7, dtype=torch.uint8, device='cuda')
        stream.fill(torch.from_numpy(uids).cuda(), torch.from_numpy(off).cuda(), d_keep)
        got = d_keep.cpu().numpy()
        assert np.array_equal(got[:off[-1]], keep[:off[-1]]), (step, np.flatnonzero(got[:off[-1]] != keep[:off[-1]])[:10])
        assert (got[off[-1]:] == 7).all()               # nothing written past the batch's entries
    # the device stream and the host generator are at the same position
    stream.sync_host(twin)
    assert [twin.random() for _ in range(50)] == [host.random() for _ in range(50)]
    lib.drb_ctx_destroy(ctx)


def test_cdae_steps_with_the_device_mt19937_stream_vs_oracle():
    """The ml-100k-sized configuration with the corruption stream replayed on the device: same losses as the oracle fed by
    CPython's own random.Random(seed) stream, step after step."""
    U, I, K, B = 300, 1682, 50, 64
    u, i, v = drb.synthetic_interactions(U, I, 30000, seed=10)
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    rng = np.random.default_rng(1)

    def g(shape, fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    w = {'W': g((I, K), I, K), 'W_': g((K, I), K, I), 'V': g((U, K), U, K), 'b': g((K,), K, K), 'b_': g((I,), I, I)}
    m = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='mt19937_device')
    m.fit(ds, epochs=0, batch_size=B, init_weights=w)
    o = CDAEOracle(w['W'], w['W_'], w['V'], w['b'], w['b_'], ds.csr(), learning_rate=1e-3)
    so, pr = PointSamplerOracle(ds.uid, ds.iid, ds.interaction, 5, 1e-3, 10), random.Random(10)
    for s in range(1, 31):
        m._step = s
        loss = m._train_step(B, 1e-3, want_loss=True, prefetch=True)
        uids = np.array([t[0] for t in so.sample(B)])
        keep = np.stack([corruption_keep_mt(pr, I, 0.2) for _ in uids])
        loss_o = float(o.step(uids, keep, 1e-3))
        assert abs(loss - loss_o) <= 1e-3 * abs(loss_o), (s, loss, loss_o)
    m2 = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='mt19937')       # host replay: identical run
    m2.fit(ds, epochs=0, batch_size=B, init_weights=w)
    m3 = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='mt19937_device')
    m3.fit(ds, epochs=0, batch_size=B, init_weights=w)
    for s in range(1, 6):
        m2._step = m3._step = s
        a, b = m2._train_step(B, 1e-3, want_loss=True), m3._train_step(B, 1e-3, want_loss=True)
        assert abs(a - b) <= 2e-6 * abs(a), (s, a, b)
