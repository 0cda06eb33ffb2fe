"""Device replay of the reference's corruption stream (rng_mode='mt19937_device').

DRecPy/Recommender/cdae.py:63-64 draws `self._rng.uniform(0, 1)` once per item for every sampled user from the one
sequential random.Random of recommender_abc.py:74.  This class keeps that stream on the GPU: the current position as a
624-word window, the jump polynomials of mt_jump.cpp for the step's segments (computed once per batch size, on the
host), and one kernel launch per step (drb_mt_keep_device) that writes the same keep bytes the host replay
(drb_cdae_corruption_keep_mt) would -- bit for bit, 1.4 ms instead of ~0.3 s per step at the ml-20m shape.
"""
import ctypes as C
from fractions import Fraction
from math import ceil

import numpy as np

from . import _lib


class DeviceMaskStream:
    def __init__(self, host_rng, n_items, corruption_level, torch, device, ctx, d_indptr, d_indices):
        self.torch, self.dev, self.ctx = torch, device, ctx
        self.n_items = int(n_items)
        self.threshold = int(ceil(Fraction(float(corruption_level)) * 2 ** 53))      # u < q  <=>  53-bit integer < this
        self.d_indptr, self.d_indices = d_indptr, d_indices
        w = np.zeros(624, np.uint32)
        _lib.check(_lib.load().drb_rng_window(host_rng.handle, _lib.np_ptr(w)))
        self.windows = [torch.from_numpy(w.view(np.int32).copy()).to(device), torch.zeros(624, dtype=torch.int32, device=device)]
        self.cur = 0
        self.tables = {}
        self.outputs_consumed = 0

    def _table(self, batch):
        """(users per CTA, polys [n_cta - 1][312] on the device, poly of the whole step) for this batch size."""
        if batch in self.tables:
            return self.tables[batch]
        lib = _lib.load()
        # enough CTAs to fill the GPU twice over, but at least ~16 regenerations of 624 words each
        ups = max(1, -(-batch // 296), -(-9984 // (2 * self.n_items)))
        ups = min(ups, 64)
        n_cta = -(-batch // ups)
        polys = None
        if n_cta > 1:
            j = _lib.vp()
            _lib.check(lib.drb_mtjump_create(ups * 2 * self.n_items, n_cta - 1, C.byref(j)))
            buf = np.zeros((n_cta - 1, 312), np.uint64)
            _lib.check(lib.drb_mtjump_polys(j, _lib.np_ptr(buf)))
            lib.drb_mtjump_destroy(j)
            polys = self.torch.from_numpy(buf.view(np.int64)).to(self.dev)
        j = _lib.vp()
        _lib.check(lib.drb_mtjump_create(batch * 2 * self.n_items, 1, C.byref(j)))
        tot = np.zeros((1, 312), np.uint64)
        _lib.check(lib.drb_mtjump_polys(j, _lib.np_ptr(tot)))
        lib.drb_mtjump_destroy(j)
        self.tables[batch] = (ups, polys, self.torch.from_numpy(tot.view(np.int64)).to(self.dev))
        return self.tables[batch]

    def fill(self, uids_dev, keep_off_dev, keep_dev):
        """Writes the keep bytes of this batch (device tensors) and advances the stream by batch * 2 * n_items outputs."""
        batch = int(uids_dev.numel())
        ups, polys, total = self._table(batch)
        w_in, w_out = self.windows[self.cur], self.windows[1 - self.cur]
        _lib.check(_lib.load().drb_mt_keep_device(
            self.ctx, _lib.t_ptr(w_in), _lib.t_ptr(w_out), _lib.t_ptr(polys), _lib.t_ptr(total), _lib.t_ptr(uids_dev),
            _lib.t_ptr(keep_off_dev), _lib.t_ptr(self.d_indptr), _lib.t_ptr(self.d_indices), _lib.t_ptr(keep_dev), batch,
            self.n_items, ups, self.threshold))
        self.cur = 1 - self.cur
        self.outputs_consumed += batch * 2 * self.n_items

    def sync_host(self, host_rng):
        """Puts the host generator at the stream's current position (e.g. before handing the model back to host code)."""
        w = self.windows[self.cur].cpu().numpy().view(np.uint32).copy()
        _lib.check(_lib.load().drb_rng_set_window(host_rng.handle, _lib.np_ptr(w)))
