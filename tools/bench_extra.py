#!/usr/bin/env python
"""Secondary measurements on one GPU (not the driver's bench contract): DMF training samples/s at config 2 and
ranked users/s at config 4 (sampled leave-1-out protocol and full-catalog top-100).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dmf_c2(steps=200):
    import torch
    import drecpy_b200 as drb
    u, i, v = drb.synthetic_interactions(6040, 3706, 1_000_000, seed=10)
    ds = drb.InteractionData(u, i, v)
    m = drb.DMF(user_factors=[64, 32], item_factors=[64, 32], seed=10, verbose=False)
    B = 256
    m.fit(ds, epochs=0, batch_size=B, learning_rate=1e-3, neg_ratio=5, reg_rate=1e-4)
    dev = torch.device('cuda')
    batches = []
    for _ in range(steps + 20):
        uu, ii, vv = m._sampler.sample_arrays(B)
        batches.append((torch.from_numpy(uu.copy()).to(dev), torch.from_numpy(ii.copy()).to(dev),
                        torch.from_numpy(m.labels_from_values(vv)).to(dev)))
    loss = torch.zeros(2, device=dev)
    for s in range(20):
        m.step_device(*batches[s], 1e-4, loss)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = m.launch_count()
    e0.record()
    for s in range(20, 20 + steps):
        m.step_device(*batches[s], 1e-4, loss)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (m.launch_count() - l0) / steps
    # end to end through fit()'s step (host sampler -> H2D -> step -> D2H loss): median over chunks of 100 consecutive
    # steps, so that a one-off host stall (a 40 ms pause was seen once per process) does not decide the number
    chunks = []
    for _ in range(max(3, steps // 100)):
        t0 = time.perf_counter()
        for _ in range(100):
            m._step += 1
            m._train_step(B, 1e-4, want_loss=True, prefetch=True)
        chunks.append((time.perf_counter() - t0) / 100)
    t_e2e = float(np.median(chunks))
    return {'dmf_c2_samples_per_s': B / (ms * 1e-3), 'dmf_c2_ms_per_step': ms, 'dmf_c2_launches_per_step': launches,
            'dmf_c2_e2e_samples_per_s': B / t_e2e, 'dmf_loss': float(loss[0])}


def eval_c4(n_users=138493, n_items=26744, nnz=20_000_000, K=200, arrays=None):
    import torch
    import drecpy_b200 as drb
    u, i, v = arrays if arrays is not None else drb.synthetic_interactions(n_users, n_items, nnz, seed=10, zipf_a=1.0)
    # config 4's test set: the reference's leave-1-out split (one held-out interaction per user with > 1 rows),
    # drecpy_b200.leave_k_out == DRecPy/Evaluation/Splits/leave_k_out.py, per-user Random(seed + idx + 1)
    t0 = time.perf_counter()
    train, test = drb.leave_k_out(drb.InteractionData(u, i, v), k=1, min_user_interactions=0, seed=10,
                                  max_concurrent_threads=16, verbose=False)
    t_split = time.perf_counter() - t0
    m = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='philox')
    m.fit(train, epochs=2, batch_size=4096)
    out = {'eval_users': int(len(test)), 'leave_k_out_s': t_split, 'leave_k_out_users_per_s': n_users / t_split}
    t0 = time.perf_counter()
    res = drb.ranking_evaluation(m, test, k=10, n_pos_interactions=1, n_neg_interactions=100,
                                 generate_negative_pairs=True, novelty=True, seed=10,
                                 metrics=[drb.HitRatio(), drb.NDCG()], verbose=False)
    dt = time.perf_counter() - t0
    out.update({'sampled_ranking_users_per_s': out['eval_users'] / dt, 'sampled_ranking_s': dt, 'metrics': res})
    uids = torch.arange(m.n_users, dtype=torch.int32, device='cuda')
    m.topk_batch(uids[:4096], 100, novelty=True, return_device=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    oi, os_, on = m.topk_batch(uids, 100, novelty=True, return_device=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.update({'full_catalog_top100_users_per_s': m.n_users / dt, 'full_catalog_s': dt})
    return out


if __name__ == '__main__':
    res = {}
    what = sys.argv[1:] or ['dmf', 'eval']
    if 'dmf' in what:
        res.update(dmf_c2())
    if 'eval' in what:
        res.update(eval_c4())
    print(json.dumps(res))
