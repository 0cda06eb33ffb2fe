"""Oracle restatement of the DMF training step and scoring (numpy fp32; test infrastructure).

PARITY UNPINNED for the arithmetic (TensorFlow absent, see oracle/__init__.py).  Follows /root/reference:
  DRecPy/Recommender/dmf.py:46-62    towers: Dense(relu) on every layer, l2(reg) on kernels, zero-init biases
  DRecPy/Recommender/dmf.py:64-86    labels (r-min)/(max-min) when use_nce; dense raw-valued rows/columns,
                                     tf.nn.l2_normalize(axis=1) when l2_norm_vectors
  DRecPy/Recommender/dmf.py:88-99    cosine of l2-normalised tower outputs, max(1e-6, .), Keras BCE
  DRecPy/Recommender/dmf.py:101-106  _predict = same forward at batch 1, rescaled to [min,max]
  DRecPy/Recommender/recommender_abc.py:314-334  reg = sum(layer.losses) = reg*sum(kernel^2); one
                                     apply_gradients per tower (Q2): t = 2(s-1)+1 (user_nn), 2(s-1)+2 (item_nn)
"""
import numpy as np
from .cdae import F, KERAS_EPS, adam_update

L2N_EPS = F(1e-12)     # tf.nn.l2_normalize epsilon


def l2_normalize(x):
    ss = (x * x).sum(axis=1, keepdims=True, dtype=F)
    inv = F(1) / np.sqrt(np.maximum(ss, L2N_EPS), dtype=F)
    return (x * inv).astype(F), ss, inv


def l2_normalize_bwd(x, ss, inv, dy):
    xhat = x * inv
    proj = (xhat * dy).sum(axis=1, keepdims=True, dtype=F)
    full = (dy - xhat * proj) * inv
    return np.where(ss >= L2N_EPS, full, dy * inv).astype(F)


class DMFOracle:
    def __init__(self, user_layers, item_layers, csr, csc, min_interaction, max_interaction, use_nce=True,
                 l2_norm_vectors=True, learning_rate=1e-3, beta1=0.9, beta2=0.999, adam_eps=1e-7,
                 adam_t='per_variable'):
        """user_layers / item_layers: list of (kernel [in,out], bias [out]); csr = user rows over items,
        csc = item rows over users, raw float values (duplicates summed)."""
        self.user_layers = [(np.array(k, F), np.array(b, F)) for k, b in user_layers]
        self.item_layers = [(np.array(k, F), np.array(b, F)) for k, b in item_layers]
        self.csr, self.csc = csr, csc
        self.n_items = self.user_layers[0][0].shape[0]
        self.n_users = self.item_layers[0][0].shape[0]
        self.min_i, self.max_i = min_interaction, max_interaction
        self.use_nce, self.l2n = use_nce, l2_norm_vectors
        self.lr, self.beta1, self.beta2, self.adam_eps = learning_rate, beta1, beta2, adam_eps
        self.adam_t = adam_t
        self.groups = [self.user_layers, self.item_layers]
        self.m = [[(np.zeros_like(k), np.zeros_like(b)) for k, b in g] for g in self.groups]
        self.v = [[(np.zeros_like(k), np.zeros_like(b)) for k, b in g] for g in self.groups]
        self.step_count = 0

    @staticmethod
    def _dense_rows(mat, ids, n):
        indptr, indices, data = mat
        out = np.zeros((len(ids), n), F)
        for r, x in enumerate(ids):
            lo, hi = indptr[x], indptr[x + 1]
            out[r, indices[lo:hi]] = data[lo:hi]
        return out

    def standardize(self, value):                       # recommender_abc.py:463-465
        return (value - self.min_i) / (self.max_i - self.min_i)

    def rescale(self, value):                           # recommender_abc.py:467-469
        return self.min_i + (self.max_i - self.min_i) * value

    def _tower_fwd(self, layers, x):
        acts = [x]
        for k, b in layers:
            x = np.maximum(x @ k + b, F(0)).astype(F)
            acts.append(x)
        return acts

    def forward(self, uids, iids):
        xu = self._dense_rows(self.csr, uids, self.n_items)          # dmf.py:76
        xi = self._dense_rows(self.csc, iids, self.n_users)          # dmf.py:77
        if self.l2n:
            xu = l2_normalize(xu)[0]
            xi = l2_normalize(xi)[0]
        au = self._tower_fwd(self.user_layers, xu)
        ai = self._tower_fwd(self.item_layers, xi)
        nu, ssu, invu = l2_normalize(au[-1])
        ni, ssi, invi = l2_normalize(ai[-1])
        c = (nu * ni).sum(axis=1, dtype=F)
        p = np.maximum(F(1e-6), c)
        return p, dict(au=au, ai=ai, nu=nu, ni=ni, ssu=ssu, ssi=ssi, invu=invu, invi=invi, c=c)

    def predict(self, uid, iid):
        return self.rescale(self.forward([uid], [iid])[0][0])

    @staticmethod
    def _tower_bwd(layers, acts, dout):
        grads = []
        d = dout
        for (k, b), a_in, a_out in zip(reversed(layers), reversed(acts[:-1]), reversed(acts[1:])):
            dpre = (d * (a_out > 0)).astype(F)
            grads.append((a_in.T @ dpre, dpre.sum(axis=0, dtype=F)))
            d = dpre @ k.T
        return list(reversed(grads))

    def grads(self, uids, iids, labels, reg_rate):
        """loss (batch + reg) and gradients [(gk, gb) per layer] for the user and the item tower."""
        B = len(uids)
        t = np.asarray(labels, dtype=np.float64).astype(F)
        p, cache = self.forward(uids, iids)
        pc = np.clip(p, KERAS_EPS, F(1) - KERAS_EPS)
        elem = -(t * np.log(pc + KERAS_EPS) + (F(1) - t) * np.log(F(1) - pc + KERAS_EPS))
        loss = F(elem.sum(dtype=np.float64) / B)
        inside = (p >= KERAS_EPS) & (p <= F(1) - KERAS_EPS)
        dp = (-(t / (pc + KERAS_EPS) - (F(1) - t) / (F(1) - pc + KERAS_EPS)) * inside / F(B)).astype(F)
        dc = (dp * (cache['c'] > F(1e-6)))[:, None].astype(F)
        dnu, dni = dc * cache['ni'], dc * cache['nu']
        dau = l2_normalize_bwd(cache['au'][-1], cache['ssu'], cache['invu'], dnu)
        dai = l2_normalize_bwd(cache['ai'][-1], cache['ssi'], cache['invi'], dni)
        gu = self._tower_bwd(self.user_layers, cache['au'], dau)
        gi = self._tower_bwd(self.item_layers, cache['ai'], dai)
        reg = F(0)
        for g in self.groups:
            for k, _ in g:
                reg += F(reg_rate) * F((k.astype(np.float64) ** 2).sum())
        out = []
        for layers, grads in zip(self.groups, (gu, gi)):
            out.append([((gk + F(2 * reg_rate) * k).astype(F), gb.astype(F)) for (k, _), (gk, gb) in zip(layers, grads)])
        return F(loss + reg), out

    def step(self, uids, iids, labels, reg_rate):
        total, all_grads = self.grads(uids, iids, labels, reg_rate)
        self.step_count += 1
        s = self.step_count
        for gidx, (layers, grads) in enumerate(zip(self.groups, all_grads)):
            tt = 2 * (s - 1) + gidx + 1 if self.adam_t == 'per_variable' else s
            for li, ((k, b), (gk, gb)) in enumerate(zip(layers, grads)):
                adam_update(k, self.m[gidx][li][0], self.v[gidx][li][0], gk, self.lr, tt,
                            self.beta1, self.beta2, self.adam_eps)
                adam_update(b, self.m[gidx][li][1], self.v[gidx][li][1], gb.astype(F), self.lr, tt,
                            self.beta1, self.beta2, self.adam_eps)
        return total
