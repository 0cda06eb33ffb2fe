"""Oracle restatement of PointSampler and the two generators it is built from (test infrastructure).

Follows /root/reference:
  DRecPy/Sampler/point_sampler.py:19-42   three independent random.Random(seed) streams
  DRecPy/Sampler/point_sampler.py:55-61   sample(): rng.uniform(0, neg_ratio+1) > 1 ? negative : positive
  DRecPy/Dataset/mem_dataset.py:119-129   positive generator (rows of a random uid, in DataFrame order,
                                          after the 'interaction >= thr' filter)
  DRecPy/Dataset/mem_dataset.py:154-163   null-pair generator (pure rejection sampling; the
                                          "existing sub-threshold pair" branch :139-147 is dead for the in-memory
                                          backend because :143 passes a DataFrame as the query and the bare
                                          except swallows the resulting error)
The RNG is CPython's own random.Random, i.e. exactly what the reference runs on.
"""
import random
import numpy as np


class PointSamplerOracle:
    def __init__(self, uid, iid, val, neg_ratio, interaction_threshold=None, seed=None):
        uid = np.asarray(uid)
        iid = np.asarray(iid)
        val = np.asarray(val)
        self.neg_ratio = neg_ratio
        self.rng = random.Random(seed)            # point_sampler.py:30
        self.null_rng = random.Random(seed)       # mem_dataset.py:135-136
        self.pos_rng = random.Random(seed)        # mem_dataset.py:113-114
        self.max_uid = int(uid.max())             # mem_dataset.py:116,149 (whole dataset, unfiltered)
        self.max_iid = int(iid.max())
        self.pairs = set(zip(uid.tolist(), iid.tolist()))   # membership over ALL rows (:161)
        keep = np.ones(len(uid), bool) if interaction_threshold is None else (val >= interaction_threshold)
        self.user_rows = {}
        for r in np.flatnonzero(keep):            # DataFrame row order
            self.user_rows.setdefault(int(uid[r]), []).append((int(iid[r]), val[r]))

    def sample_negative(self):
        while True:
            u = self.null_rng.randint(0, self.max_uid)
            i = self.null_rng.randint(0, self.max_iid)
            if (u, i) not in self.pairs:
                return u, i, 0

    def sample_positive(self):
        while True:
            u = self.pos_rng.randint(0, self.max_uid)
            rows = self.user_rows.get(u)
            if not rows:
                continue
            j = self.pos_rng.randint(0, len(rows) - 1)
            return u, rows[j][0], rows[j][1]

    def sample(self, n=16):
        out = []
        while len(out) != n:
            null_pair = self.rng.uniform(0, self.neg_ratio + 1) > 1
            out.append(self.sample_negative() if null_pair else self.sample_positive())
        return out


class PointSamplerOracleCSR(PointSamplerOracle):
    """The same three streams and the same draws as PointSamplerOracle, with array storage instead of a Python set of
    all (uid, iid) pairs and a dict of row lists, so that it can be built for the 20 M-row shape in seconds (the CPU
    arm of bench.py samples with it).  Membership of mem_dataset.py:161 = binary search in the user's sorted item
    list; positives = the user's rows in DataFrame order (mem_dataset.py:119-129)."""

    def __init__(self, uid, iid, val, neg_ratio, interaction_threshold=None, seed=None):
        uid = np.asarray(uid).astype(np.int64)
        iid = np.asarray(iid).astype(np.int64)
        val = np.asarray(val)
        self.neg_ratio = neg_ratio
        self.rng = random.Random(seed)
        self.null_rng = random.Random(seed)
        self.pos_rng = random.Random(seed)
        self.max_uid = int(uid.max())
        self.max_iid = int(iid.max())
        n_u = self.max_uid + 1
        order = np.lexsort((iid, uid))                                   # all rows: by user, items ascending
        self.all_indptr = np.concatenate(([0], np.cumsum(np.bincount(uid, minlength=n_u))))
        self.all_iid = iid[order]
        sel = np.arange(len(uid)) if interaction_threshold is None else np.flatnonzero(val >= interaction_threshold)
        o2 = sel[np.argsort(uid[sel], kind='stable')]                    # filtered rows: by user, DataFrame order kept
        self.pos_indptr = np.concatenate(([0], np.cumsum(np.bincount(uid[sel], minlength=n_u))))
        self.pos_iid = iid[o2]
        self.pos_val = val[o2]

    def sample_negative(self):
        while True:
            u = self.null_rng.randint(0, self.max_uid)
            i = self.null_rng.randint(0, self.max_iid)
            lo, hi = self.all_indptr[u], self.all_indptr[u + 1]
            p = lo + np.searchsorted(self.all_iid[lo:hi], i)
            if p == hi or self.all_iid[p] != i:
                return u, i, 0

    def sample_positive(self):
        while True:
            u = self.pos_rng.randint(0, self.max_uid)
            lo, hi = int(self.pos_indptr[u]), int(self.pos_indptr[u + 1])
            if hi == lo:
                continue
            j = self.pos_rng.randint(0, hi - lo - 1)
            return u, int(self.pos_iid[lo + j]), self.pos_val[lo + j]
