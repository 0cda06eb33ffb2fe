"""Cross-check the oracle's hand-derived CDAE / DMF gradients against torch.autograd in float64, using the
LITERAL Keras-2 forms (the (B,B,I) loss broadcast of SURVEY.md Q1 for CDAE).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import dataset as ods
from oracle.cdae import CDAEOracle
from oracle.dmf import DMFOracle

EPS = 1e-7


def _toy(U=17, I=23, nnz=120, seed=0):
    rng = np.random.default_rng(seed)
    pairs = rng.choice(U * I, nnz, replace=False)
    uid, iid = (pairs // I).astype(np.int32), (pairs % I).astype(np.int32)
    val = rng.integers(1, 6, nnz).astype(np.float64)
    return uid, iid, val


def keras_bce(y_true, y_pred):
    p = torch.clamp(y_pred, EPS, 1 - EPS)
    bce = y_true * torch.log(p + EPS) + (1 - y_true) * torch.log(1 - p + EPS)
    return (-bce).mean(dim=-1).mean()


@pytest.mark.parametrize('loss', ['bce', 'mse'])
def test_cdae_grads_vs_autograd_literal_broadcast(loss):
    U, I, K, B, q, reg = 17, 23, 6, 5, 0.2, 1e-2
    uid, iid, val = _toy(U, I)
    csr = ods.build_csr(uid, iid, val, U, I)
    rng = np.random.default_rng(1)
    W, W_, V = rng.normal(0, .3, (I, K)), rng.normal(0, .3, (K, I)), rng.normal(0, .3, (U, K))
    b, b_ = rng.normal(0, .3, K), rng.normal(0, .3, I)
    orc = CDAEOracle(W, W_, V, b, b_, csr, corruption_level=q, loss=loss)
    uids = np.array([3, 9, 3, 0, 16])
    keep = rng.random((B, I)) >= q
    total, grads = orc.grads(uids, keep, reg)

    tW, tW_, tV, tb, tb_ = [torch.tensor(np.array(x, np.float32).astype(np.float64), requires_grad=True)
                            for x in (W, W_, V, b, b_)]
    y = torch.tensor(orc.desired(uids).astype(np.float64))
    x = y * torch.tensor(keep.astype(np.float64)) / (1 - q)
    preds = []
    for r in range(B):                                          # cdae.py:50-57: list of (1, I) tensors
        h = torch.sigmoid(x[r:r + 1] @ tW + tV[uids[r]] + tb)
        preds.append(torch.sigmoid(h @ tW_ + tb_))
    y_pred = torch.stack(preds)                                 # (B, 1, I)
    if loss == 'bce':
        L = keras_bce(y, y_pred)                                # (B,I) vs (B,1,I) -> (B,B,I)
    else:
        L = ((y_pred - y) ** 2).mean(dim=-1).mean()
    L = L + sum(0.5 * (t ** 2).sum() for t in (tW, tW_, tV)) * reg / B
    L.backward()
    assert abs(float(total) - float(L.detach())) < 1e-5 * abs(float(L.detach()))
    for g, t in zip(grads, (tW, tW_, tV, tb, tb_)):
        ref = t.grad.numpy()
        assert np.allclose(g, ref, rtol=2e-4, atol=1e-7), np.abs(g - ref).max()


def test_cdae_per_user_labels_differ_from_batch_mean():
    U, I, K = 17, 23, 6
    uid, iid, val = _toy(U, I)
    csr = ods.build_csr(uid, iid, val, U, I)
    rng = np.random.default_rng(2)
    args = [rng.normal(0, .3, s) for s in ((I, K), (K, I), (U, K), (K,), (I,))]
    a = CDAEOracle(*args, csr, label_mode='batch_mean')
    b = CDAEOracle(*args, csr, label_mode='per_user')
    uids = np.array([1, 2, 3])
    keep = np.ones((3, I), bool)
    assert abs(float(a.grads(uids, keep, 0.0)[0]) - float(b.grads(uids, keep, 0.0)[0])) > 1e-4


def test_dmf_grads_vs_autograd():
    U, I, B, reg = 19, 29, 7, 1e-3
    uid, iid, val = _toy(U, I, 200, seed=3)
    csr = ods.build_csr(uid, iid, val, U, I)
    csc = ods.build_csr(iid, uid, val, I, U)
    rng = np.random.default_rng(4)
    ul = [(rng.normal(0, .4, (I, 8)), rng.normal(0, .1, 8)), (rng.normal(0, .4, (8, 4)), rng.normal(0, .1, 4))]
    il = [(rng.normal(0, .4, (U, 8)), rng.normal(0, .1, 8)), (rng.normal(0, .4, (8, 4)), rng.normal(0, .1, 4))]
    orc = DMFOracle(ul, il, csr, csc, 0, 5)
    uids = rng.integers(0, U, B)
    iids = rng.integers(0, I, B)
    labels = rng.integers(0, 6, B) / 5.0
    total, grads = orc.grads(uids, iids, labels, reg)

    def tt(a):
        return torch.tensor(np.array(a, np.float32).astype(np.float64), requires_grad=True)
    tul = [(tt(k), tt(b)) for k, b in ul]
    til = [(tt(k), tt(b)) for k, b in il]
    xu = torch.tensor(orc._dense_rows(csr, uids, I).astype(np.float64))
    xi = torch.tensor(orc._dense_rows(csc, iids, U).astype(np.float64))

    def l2n(x):
        return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12))
    a, e = l2n(xu), l2n(xi)
    for k, b in tul:
        a = torch.relu(a @ k + b)
    for k, b in til:
        e = torch.relu(e @ k + b)
    c = (l2n(a) * l2n(e)).sum(1)
    p = torch.maximum(torch.tensor(1e-6, dtype=torch.float64), c)
    L = keras_bce(torch.tensor(labels), p) + reg * sum((k ** 2).sum() for k, _ in tul + til)
    L.backward()
    assert abs(float(total) - float(L.detach())) < 1e-5 * abs(float(L.detach()))
    for g_layers, t_layers in zip(grads, (tul, til)):
        for (gk, gb), (tk, tb) in zip(g_layers, t_layers):
            for g, ref in ((gk, tk.grad.numpy()), (gb, tb.grad.numpy())):
                # fp32 cancellation noise scales with the largest gradient entry (p clamps at 1e-6 -> 1/p ~ 1e6)
                assert np.allclose(g, ref, rtol=2e-4, atol=2e-6 * max(1.0, np.abs(ref).max()))


def test_adam_matches_torch_adam_with_keras_epsilon_placement():
    """Keras: w -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps)  (eps outside the bias-corrected sqrt)."""
    from oracle.cdae import adam_update
    rng = np.random.default_rng(0)
    w = rng.normal(size=10).astype(np.float32)
    m = np.zeros_like(w)
    v = np.zeros_like(w)
    w64, m64, v64 = w.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for t in range(1, 6):
        g = rng.normal(size=10).astype(np.float32)
        adam_update(w, m, v, g, 1e-3, t)
        m64 = 0.9 * m64 + 0.1 * g
        v64 = 0.999 * v64 + 0.001 * g.astype(np.float64) ** 2
        w64 -= 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m64 / (np.sqrt(v64) + 1e-7)
    assert np.allclose(w, w64, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('loss', ['bce', 'mse'])
def test_cdae_sampled_output_grads_vs_autograd(loss):
    """The sampled-output extension (configs[4]): positives + n_groups x neg_per_group drawn items per sampled user,
    per-user labels, loss / (B * n_neg).  Hand-derived gradients of oracle.cdae.CDAESampledOracle vs autograd (fp64)."""
    from oracle.cdae import CDAESampledOracle, sampled_negatives
    U, I, K, B, q, reg, G, npg, step, seed = 17, 23, 6, 5, 0.2, 1e-2, 2, 4, 3, 10
    uid, iid, val = _toy(U, I)
    csr = ods.build_csr(uid, iid, val, U, I)
    rng = np.random.default_rng(1)
    W, W_, V = rng.normal(0, .3, (I, K)), rng.normal(0, .3, (K, I)), rng.normal(0, .3, (U, K))
    b, b_ = rng.normal(0, .3, K), rng.normal(0, .3, I)
    orc = CDAESampledOracle(W, W_, V, b, b_, csr, corruption_level=q, loss=loss, n_groups=G, neg_per_group=npg, seed=seed)
    uids = np.array([3, 9, 3, 0, 16])
    keep = rng.random((B, I)) >= q
    total, grads = orc.grads_sampled(uids, keep, reg, step)

    tW, tW_, tV, tb, tb_ = [torch.tensor(np.array(x, np.float32).astype(np.float64), requires_grad=True)
                            for x in (W, W_, V, b, b_)]
    y = torch.tensor(orc.desired(uids).astype(np.float64))
    x = y * torch.tensor(keep.astype(np.float64)) / (1 - q)
    L = 0
    for r in range(B):
        h = torch.sigmoid(x[r:r + 1] @ tW + tV[uids[r]] + tb)
        neg = sampled_negatives(I, G, npg, r, step, seed)
        assert len(neg) == G * npg and (neg[:npg] < 12).all() and (neg[npg:] >= 12).all()   # group ranges [0,12), [12,23)
        items = np.concatenate([orc.positives(uids[r]), neg])
        p = torch.sigmoid(h @ tW_[:, items] + tb_[items])[0]
        t = y[r, items]
        if loss == 'bce':
            pc = torch.clamp(p, EPS, 1 - EPS)
            L = L - (t * torch.log(pc + EPS) + (1 - t) * torch.log(1 - pc + EPS)).sum()
        else:
            L = L + ((p - t) ** 2).sum()
    L = L / (B * G * npg) + sum(0.5 * (t_ ** 2).sum() for t_ in (tW, tW_, tV)) * reg / B
    L.backward()
    assert abs(float(total) - float(L.detach())) < 1e-5 * abs(float(L.detach()))
    for g, t_ in zip(grads, (tW, tW_, tV, tb, tb_)):
        ref = t_.grad.numpy()
        assert np.allclose(g, ref, rtol=2e-4, atol=1e-7), np.abs(g - ref).max()
