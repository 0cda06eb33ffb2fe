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
