"""PointSampler: positive / negative (uid, iid, value) pair sampling, bit-exact with the reference.

Mirrors DRecPy/Sampler/point_sampler.py:19-96 (same constructor, sample / sample_one / sample_negative /
sample_positive); the three MT19937 streams and the two generators of DRecPy/Dataset/mem_dataset.py:101-163 are
replayed by the host C++ sampler in libdrb (drb_sampler_*), which replaces the reference's O(nnz) pandas scan per
draw with a CSR membership test.
"""
import ctypes as C
import random

import numpy as np

from . import _lib
from .dataset import InteractionData


class PointSampler:
    def __init__(self, interaction_dataset, neg_ratio, interaction_threshold=None, seed=None):
        assert interaction_dataset is not None, 'An interaction dataset instance is required.'
        assert neg_ratio is not None, 'A neg_ratio value is required.'
        data = InteractionData.from_dataset(interaction_dataset)
        assert data.has_internal_ids or getattr(interaction_dataset, 'has_internal_ids', False), \
            'The provided interaction dataset instance does not have internal ids assigned.'
        data.assign_internal_ids()
        assert len(data) > 0, 'No records were found to sample from.'
        self.interaction_dataset = interaction_dataset
        self.neg_ratio = neg_ratio
        self.interaction_threshold = interaction_threshold
        if seed is None:   # the reference falls back to OS entropy; draw a seed once so the run is replayable
            seed = random.SystemRandom().getrandbits(63)
        self.seed = seed
        self._int_values = bool(np.issubdtype(data.interaction.dtype, np.integer))
        self._pos = data.rows_by_user(interaction_threshold)         # (indptr, iid, val) in DataFrame order
        all_indptr, all_iid, _ = data.csr()
        self._all = (np.ascontiguousarray(all_indptr), np.ascontiguousarray(all_iid))
        self._h = _lib.vp()
        lib = _lib.load()
        _lib.check(lib.drb_sampler_create(int(data.uid.max()), int(data.iid.max()),
                                          _lib.np_ptr(self._pos[0]), _lib.np_ptr(self._pos[1]),
                                          _lib.np_ptr(self._pos[2]), _lib.np_ptr(self._all[0]),
                                          _lib.np_ptr(self._all[1]), float(neg_ratio), abs(int(seed)),
                                          C.byref(self._h)))

    def __del__(self):
        if getattr(self, '_h', None) and _lib._lib is not None:
            _lib._lib.drb_sampler_destroy(self._h)
            self._h = None

    def sample_arrays(self, n, out=None):
        """n triples as arrays (uid int32, iid int32, value float64) -- the native form fit() consumes."""
        if out is None:
            out = (np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64))
        _lib.check(_lib.load().drb_sampler_sample(self._h, n, _lib.np_ptr(out[0]), _lib.np_ptr(out[1]),
                                                  _lib.np_ptr(out[2])))
        return out

    def sample(self, n=16):
        u, i, v = self.sample_arrays(n)
        vals = [int(c) if self._int_values else c for c in v.tolist()]
        return list(zip(u.tolist(), i.tolist(), vals))

    def sample_one(self):
        return self.sample(n=1)[0]

    def getstate(self):
        st = np.zeros(3 * 625, np.uint32)
        _lib.check(_lib.load().drb_sampler_getstate(self._h, _lib.np_ptr(st)))
        return st

    def setstate(self, st):
        st = np.ascontiguousarray(st, np.uint32)
        _lib.check(_lib.load().drb_sampler_setstate(self._h, _lib.np_ptr(st)))
