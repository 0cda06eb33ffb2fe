#!/usr/bin/env python
"""Per-kernel times of the full-catalog top-k at the C3 shape (profiling aid; DRB_SCORE_DEBUG experiments).
usage: python tools/prof_topk.py [n_users] [epochs]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import drecpy_b200 as drb
from drecpy_b200 import _lib

n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
u, i, v = drb.synthetic_interactions(138493, 26744, 20_000_000, seed=10, zipf_a=1.0)
ds = drb.InteractionData(u, i, v)
m = drb.CDAE(hidden_factors=200, seed=10, verbose=False, rng_mode='philox')
m.fit(ds, epochs=epochs, batch_size=4096, score_batch=int(os.environ.get('DRB_BENCH_SCORE_BATCH', 18944)))
uids = torch.arange(n_users, dtype=torch.int32, device='cuda')
lib = _lib.load()
for dbg in os.environ.get('DBG_LIST', '0').split(','):
    os.environ['DRB_SCORE_DEBUG'] = dbg
    m.topk_batch(uids[:min(n_users, 18944)], 100, novelty=True, return_device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.topk_batch(uids, 100, novelty=True, return_device=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    _lib.check(lib.drb_ctx_profile_enable(m._ctx, 1))
    o = m.topk_batch(uids, 100, novelty=True, return_device=True)
    prof = _lib.profile_read(m._ctx)
    _lib.check(lib.drb_ctx_profile_enable(m._ctx, 0))
    print(json.dumps({'debug': dbg, 'users': n_users, 'ms': round(ms, 3), 'users_per_s': round(n_users / ms * 1e3),
                      'kernels_ms': {k: round(v[0], 3) for k, v in prof.items()}}), flush=True)
