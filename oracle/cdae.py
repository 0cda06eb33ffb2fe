"""Oracle restatement of the CDAE training step and scoring (numpy fp32; test infrastructure).

PARITY UNPINNED for the arithmetic (TensorFlow absent, see oracle/__init__.py).  Follows /root/reference:
  DRecPy/Recommender/cdae.py:59-76    _reconstruct_for_training / _reconstruct_for_predictions / _reconstruct
  DRecPy/Recommender/cdae.py:78-82    _compute_batch_loss (Keras-2 BinaryCrossentropy / MeanSquaredError on a
                                      list of (1,I) predictions vs a list of I-lists -> (B,B,I) broadcast, i.e.
                                      batch-mean labels) and _compute_reg_loss
  DRecPy/Recommender/cdae.py:90-103   _rank (heapq.nlargest over (score, iid) tuples)
  DRecPy/Recommender/recommender_abc.py:186-205,328-334   step loop; one apply_gradients per variable (Q2)
Keras-2 definitions restated: backend.binary_crossentropy (clip to [eps,1-eps], log(p+eps), eps=1e-7),
Adam (lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= lr_t*m/(sqrt(v)+eps)).
"""
import heapq
import random
import numpy as np

F = np.float32
KERAS_EPS = F(1e-7)


def sigmoid(x):
    return (F(1) / (F(1) + np.exp(-x, dtype=F))).astype(F)


def corruption_keep_mt(rng: random.Random, n_items, q):
    """cdae.py:63-64 -- one rng.uniform(0,1) draw per item, in item order, zeros included (Q4).
    Returns a bool vector: True = kept (draw >= q)."""
    return np.array([not (rng.uniform(0, 1) < q) for _ in range(n_items)], dtype=bool)


def adam_update(w, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam dense update (ResourceApplyAdam), in fp32, t = optimizer.iterations + 1."""
    b1, b2 = F(beta1), F(beta2)
    lr_t = F(lr) * np.sqrt(F(1) - np.power(b2, F(t), dtype=F), dtype=F) / (F(1) - np.power(b1, F(t), dtype=F))
    m += (g - m) * (F(1) - b1)
    v += (g * g - v) * (F(1) - b2)
    w -= lr_t * m / (np.sqrt(v, dtype=F) + F(eps))


class CDAEOracle:
    """State = the five reference variables in registration order [W, W_, V, b, b_] (cdae.py:43)."""

    def __init__(self, W, W_, V, b, b_, csr, interaction_threshold=1e-3, corruption_level=0.2, loss='bce',
                 label_mode='batch_mean', adam_t='per_variable', learning_rate=1e-3,
                 beta1=0.9, beta2=0.999, adam_eps=1e-7):
        self.W = np.array(W, F)          # [I, K]
        self.W_ = np.array(W_, F)        # [K, I]
        self.V = np.array(V, F)          # [U, K]
        self.b = np.array(b, F)          # [K]
        self.b_ = np.array(b_, F)        # [I]
        self.indptr, self.indices, self.data = csr
        self.thr = interaction_threshold
        self.q = corruption_level
        self.loss = loss
        self.label_mode = label_mode
        self.adam_t = adam_t
        self.lr, self.beta1, self.beta2, self.adam_eps = learning_rate, beta1, beta2, adam_eps
        self.vars = [self.W, self.W_, self.V, self.b, self.b_]
        self.m = [np.zeros_like(x) for x in self.vars]
        self.v = [np.zeros_like(x) for x in self.vars]
        self.step_count = 0
        self.n_items = self.W.shape[0]

    # -- cdae.py:61-62
    def positives(self, uid):
        lo, hi = self.indptr[uid], self.indptr[uid + 1]
        sel = self.data[lo:hi] >= self.thr
        return self.indices[lo:hi][sel]

    def desired(self, uids):
        y = np.zeros((len(uids), self.n_items), F)
        for r, u in enumerate(uids):
            y[r, self.positives(u)] = 1
        return y

    # -- cdae.py:73-76
    def reconstruct(self, x, uids):
        h = sigmoid(x @ self.W + self.V[uids] + self.b)
        p = sigmoid(h @ self.W_ + self.b_)
        return h, p

    # -- cdae.py:67-71, :84-88
    def predict(self, uid):
        y = self.desired([uid])
        return self.reconstruct(y, [uid])[1][0]

    def hidden(self, uids):
        return self.reconstruct(self.desired(uids), list(uids))[0]

    # -- cdae.py:90-103
    def rank(self, uid, iids, n, novelty):
        p = self.predict(uid)
        cand = set(int(i) for i in iids)
        if novelty:
            lo, hi = self.indptr[uid], self.indptr[uid + 1]
            cand -= set(int(i) for i in self.indices[lo:hi])     # every stored row of the user (cdae.py:93-98)
        return heapq.nlargest(n, [(p[i], i) for i in range(self.n_items) if i in cand])

    def loss_and_grad(self, p, y):
        """Returns (loss, dL/dp).  label_mode batch_mean == the (B,B,I) Keras broadcast (Q1)."""
        B, I = p.shape
        t = y.mean(axis=0, dtype=F)[None, :] if self.label_mode == 'batch_mean' else y
        inv = F(1.0 / (B * I))
        if self.loss == 'bce':
            pc = np.clip(p, KERAS_EPS, F(1) - KERAS_EPS)
            elem = -(t * np.log(pc + KERAS_EPS) + (F(1) - t) * np.log(F(1) - pc + KERAS_EPS))
            inside = (p >= KERAS_EPS) & (p <= F(1) - KERAS_EPS)
            dp = -(t / (pc + KERAS_EPS) - (F(1) - t) / (F(1) - pc + KERAS_EPS)) * inside * inv
            return F(elem.sum(dtype=np.float64) * inv), dp.astype(F)
        if self.loss == 'mse':
            if self.label_mode == 'batch_mean':
                elem = p * p - F(2) * p * t + t          # y^2 == y
            else:
                elem = (p - t) ** 2
            dp = F(2) * (p - t) * inv
            return F(elem.sum(dtype=np.float64) * inv), dp.astype(F)
        raise ValueError(self.loss)

    def grads(self, uids, keep, reg_rate):
        """loss (batch + reg, recommender_abc.py:200-201) and its gradient w.r.t. [W, W_, V, b, b_]."""
        uids = np.asarray(uids)
        B = len(uids)
        s = F(1.0 / (1.0 - self.q))
        y = self.desired(uids)
        x = (y * keep * s).astype(F)
        h, p = self.reconstruct(x, uids)
        loss, dp = self.loss_and_grad(p, y)
        c = F(reg_rate / B)                                              # cdae.py:82
        reg = c * F(0.5) * sum(F((w.astype(np.float64) ** 2).sum()) for w in (self.W, self.W_, self.V))
        dz2 = dp * p * (F(1) - p)
        gW_ = h.T @ dz2 + c * self.W_
        gb_ = dz2.sum(axis=0, dtype=F)
        dh = dz2 @ self.W_.T
        dz1 = dh * h * (F(1) - h)
        gb = dz1.sum(axis=0, dtype=F)
        gV = c * self.V
        np.add.at(gV, uids, dz1)
        gW = x.T @ dz1 + c * self.W
        return F(loss + reg), [gW, gW_, gV, gb, gb_]

    def step(self, uids, keep, reg_rate):
        """One optimizer step (recommender_abc.py:190-205).  keep: bool [B, I] (True = kept, cdae.py:63).
        Returns the reported loss = batch loss + reg loss, evaluated at the pre-update weights."""
        total, grads = self.grads(uids, keep, reg_rate)
        self.step_count += 1
        sidx = self.step_count
        for j, (w, m, v, g) in enumerate(zip(self.vars, self.m, self.v, grads)):
            t = 5 * (sidx - 1) + j + 1 if self.adam_t == 'per_variable' else sidx   # Q2
            adam_update(w, m, v, g.astype(F), self.lr, t, self.beta1, self.beta2, self.adam_eps)
        return total


# ---------------------------------------------------------------------------------------------------------------------
# Sampled-output extension (NOT a reference behaviour; parity unpinned by the reference, pinned against this restatement).
# BASELINE.json configs[4] (10 M users x 1 M items) cannot afford the reference's dense output layer (cdae.py:76 scores
# all I items for every sampled user: 1.5 GFLOP per user at K = 256).  The sampled form scores, for sampled user b, its
# positives N(u_b) and n_groups x neg_per_group uniformly drawn items (group g = a contiguous item range; one group per
# item shard when the weights are sharded), with per-user labels (the CDAE paper's form) and the loss normalised by
# B * n_groups * neg_per_group.  Everything else of the step (hidden layer, dense Adam + L2, per-variable step counter)
# is the reference's.
def sampled_negatives(n_items, n_groups, neg_per_group, slot, step, seed):
    """The negative items of batch slot `slot` at optimizer step `step`: philox4x32-10 draws, counter = (draw, slot,
    step_lo, step_hi), key = (seed_lo ^ 0x9E3779B9 * (group + 1), seed_hi); item = group_lo + x mod group_size."""
    from .philox import philox_first
    per = -(-n_items // n_groups)
    out = []
    d = np.arange(neg_per_group, dtype=np.uint32)
    for g in range(n_groups):
        lo, hi = min(n_items, g * per), min(n_items, (g + 1) * per)
        k0 = (seed & 0xFFFFFFFF) ^ ((0x9E3779B9 * (g + 1)) & 0xFFFFFFFF)
        x = philox_first(d, np.full_like(d, slot), np.full_like(d, step & 0xFFFFFFFF), np.full_like(d, step >> 32),
                         k0, seed >> 32)
        out.append(lo + (x.astype(np.int64) % max(hi - lo, 1)))
    return np.concatenate(out)


class CDAESampledOracle(CDAEOracle):
    def __init__(self, *args, n_groups=1, neg_per_group=64, seed=10, **kw):
        super().__init__(*args, **kw)
        self.n_groups, self.neg_per_group, self.seed = n_groups, neg_per_group, seed

    def grads_sampled(self, uids, keep, reg_rate, step):
        uids = np.asarray(uids)
        B, I = len(uids), self.n_items
        s = F(1.0 / (1.0 - self.q))
        y = self.desired(uids)
        x = (y * keep * s).astype(F)
        h = sigmoid(x @ self.W + self.V[uids] + self.b)
        inv = F(1.0 / (B * self.n_groups * self.neg_per_group))
        gW_ = np.zeros_like(self.W_)
        gb_ = np.zeros_like(self.b_)
        dh = np.zeros_like(h)
        loss = 0.0
        for b, u in enumerate(uids):
            pos = self.positives(u)
            neg = sampled_negatives(I, self.n_groups, self.neg_per_group, b, step, self.seed)
            items = np.concatenate([pos, neg]).astype(np.int64)
            t = np.concatenate([np.ones(len(pos), F), y[b, neg]])            # a drawn positive is labelled positive
            z = (h[b] @ self.W_[:, items] + self.b_[items]).astype(F)
            p = sigmoid(z)
            if self.loss == 'bce':
                pc = np.clip(p, KERAS_EPS, F(1) - KERAS_EPS)
                elem = -(t * np.log(pc + KERAS_EPS) + (F(1) - t) * np.log(F(1) - pc + KERAS_EPS))
                inside = (p >= KERAS_EPS) & (p <= F(1) - KERAS_EPS)
                dp = -(t / (pc + KERAS_EPS) - (F(1) - t) / (F(1) - pc + KERAS_EPS)) * inside * inv
            else:
                elem = (p - t) ** 2
                dp = F(2) * (p - t) * inv
            loss += float(elem.sum(dtype=np.float64))
            dz = (dp * p * (F(1) - p)).astype(F)
            np.add.at(gW_.T, items, dz[:, None] * h[b][None, :])
            np.add.at(gb_, items, dz)
            dh[b] = dz @ self.W_[:, items].T
        c = F(reg_rate / B)
        reg = c * F(0.5) * sum(F((w.astype(np.float64) ** 2).sum()) for w in (self.W, self.W_, self.V))
        dz1 = dh * h * (F(1) - h)
        gV = c * self.V
        np.add.at(gV, uids, dz1)
        gW = x.T @ dz1 + c * self.W
        return F(loss * float(inv) + reg), [gW, gW_ + c * self.W_, gV, dz1.sum(axis=0, dtype=F), gb_]

    def step_sampled(self, uids, keep, reg_rate, step):
        total, grads = self.grads_sampled(uids, keep, reg_rate, step)
        self.step_count += 1
        sidx = self.step_count
        for j, (w, m, v, g) in enumerate(zip(self.vars, self.m, self.v, grads)):
            t = 5 * (sidx - 1) + j + 1 if self.adam_t == 'per_variable' else sidx
            adam_update(w, m, v, g.astype(F), self.lr, t, self.beta1, self.beta2, self.adam_eps)
        return total
