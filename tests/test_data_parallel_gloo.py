"""Host logic of the data-parallel step on CPU with world_size-2 gloo: batch sharding, label-histogram and
gradient all-reduce reproduce the single-process step (checked with the oracle), global loss assembly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import drecpy_b200 as drb
from drecpy_b200.parallel import DataParallel, shard_slices
from oracle import dataset as ods
from oracle.cdae import CDAEOracle


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    U, I, K, Bg = 40, 57, 8, 12
    rng = np.random.default_rng(0)
    pairs = rng.choice(U * I, 500, replace=False)
    uid, iid = (pairs // I).astype(np.int32), (pairs % I).astype(np.int32)
    val = rng.integers(1, 6, 500).astype(np.float64)
    csr = ods.build_csr(uid, iid, val, U, I)
    w = [rng.normal(0, .3, s).astype(np.float32) for s in ((I, K), (K, I), (U, K), (K,), (I,))]
    uids = rng.integers(0, U, Bg)
    keep = rng.random((Bg, I)) >= 0.2
    return U, I, K, Bg, csr, w, uids, keep


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dp = DataParallel(dist)
    U, I, K, Bg, csr, w, uids, keep = _problem()
    lo, hi = dp.shard(Bg)
    assert (lo, hi) == shard_slices(Bg, world)[rank]
    o = CDAEOracle(*w, csr)
    # PREP: local label histogram -> all-reduce -> global batch-mean labels
    y_local = o.desired(uids[lo:hi])
    count = torch.from_numpy(y_local.sum(0))
    dp.all_reduce_sum(count)
    ybar = (count.numpy() / np.float32(Bg)).astype(np.float32)
    # GRADS: local forward/backward against the GLOBAL mean labels and 1/(Bg*I), L2 scaled by 1/Bg applied once
    s = np.float32(1.0 / 0.8)
    x = (y_local * keep[lo:hi] * s).astype(np.float32)
    h, p = o.reconstruct(x, uids[lo:hi])
    eps = np.float32(1e-7)
    pc = np.clip(p, eps, 1 - eps)
    inv = np.float32(1.0 / (Bg * I))
    loss_local = float((-(ybar * np.log(pc + eps) + (1 - ybar) * np.log(1 - pc + eps))).sum() * inv)
    dp_ = -(ybar / (pc + eps) - (1 - ybar) / (1 - pc + eps)) * inv
    dz2 = dp_ * p * (1 - p)
    gW_ = h.T @ dz2
    dz1 = (dz2 @ o.W_.T) * h * (1 - h)
    gV = np.zeros_like(o.V)
    np.add.at(gV, uids[lo:hi], dz1)
    gW = x.T @ dz1
    flat = torch.from_numpy(np.concatenate([gW.ravel(), gW_.ravel(), gV.ravel(), dz1.sum(0), dz2.sum(0)]))
    dp.all_reduce_sum(flat)
    reg = 0.01
    loss2 = torch.tensor([loss_local + 1.0, loss_local])      # [batch term + "reg" (=1.0 here), batch term]
    total = float(dp.global_loss(loss2))
    if rank == 0:
        np.save(out, np.concatenate([flat.numpy(), [total]]))
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step(tmp_path):
    out = str(tmp_path / 'dp.npy')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    U, I, K, Bg, csr, w, uids, keep = _problem()
    o = CDAEOracle(*w, csr)
    total, grads = o.grads(uids, keep, 0.0)                    # single process, global batch, no L2
    want = np.concatenate([grads[0].ravel(), grads[1].ravel(), grads[2].ravel(), grads[3], grads[4]])
    assert np.allclose(got[:-1], want, rtol=2e-4, atol=1e-7)
    assert abs(got[-1] - (float(total) + 1.0)) < 1e-5


def test_replay_sampler_shards_partition_the_global_batch():
    u, i, v = drb.synthetic_interactions(80, 120, 2000, seed=2)
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    single = drb.PointSampler(ds, 5, 1e-3, 10).sample_arrays(24)[0]
    parts = []
    for lo, hi in shard_slices(24, 4):
        parts.append(drb.PointSampler(ds, 5, 1e-3, 10).sample_arrays(24)[0][lo:hi])   # every rank replays the stream
    assert np.array_equal(np.concatenate(parts), single)
