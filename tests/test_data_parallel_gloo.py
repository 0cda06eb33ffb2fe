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


# ------------------------------------------------------------------------------------------------------------------
# The schedule itself: drecpy_b200.parallel.data_parallel_step (the function the GPU path runs) driven on CPU by an
# oracle-backed stand-in for the native phases, 2 ranks over gloo, against the single-process oracle step on the global
# batch.  This checks what is reduced, gathered and updated, in which order, and that the three update launches see
# final gradients.
class OraclePhases:
    """drb_cdae_step_phases restated on numpy for one rank's slice of the batch (same buffers, same cut points)."""

    def __init__(self, w, csr, uids, keep, gbatch, lr=1e-3, reg=1e-3, step=1):
        from oracle.cdae import adam_update, sigmoid
        self.adam_update, self.sigmoid = adam_update, sigmoid
        self.o = CDAEOracle(*w, csr, learning_rate=lr)
        self.uids, self.keep, self.gbatch, self.reg, self.step = np.asarray(uids), keep, gbatch, reg, step
        I, K = self.o.W.shape
        U = self.o.V.shape[0]
        self.I, self.K, self.U = I, K, U
        self.off_w, self.off_v = I * K, 2 * I * K + K + I            # arena: [W2T | W | b | b2 | V]
        self.grads = torch.zeros(self.off_v + U * K, dtype=torch.float32)
        self.label_count = torch.zeros(I, dtype=torch.float32)
        self.dz1 = torch.zeros((len(uids), K), dtype=torch.float32)
        self.loss = None
        self.log = []

    def __call__(self, mask):
        o, g = self.o, self.grads.numpy()
        I, K = self.I, self.K
        self.log.append(mask)
        if mask == 1:                                                  # PREP
            g[:] = 0
            self.y = o.desired(self.uids)
            self.label_count.numpy()[:] = self.y.sum(0)
        elif mask == 2:                                                # GRADS_A: hidden layer, needs no labels
            self.x = (self.y * self.keep * np.float32(1 / 0.8)).astype(np.float32)
            self.h, self.p = o.reconstruct(self.x, self.uids)
        elif mask == 8:                                                # GRADS_B: needs the GLOBAL label histogram
            eps = np.float32(1e-7)
            t = (self.label_count.numpy() / np.float32(self.gbatch))[None, :]
            pc = np.clip(self.p, eps, 1 - eps)
            inv = np.float32(1.0 / (self.gbatch * I))
            self.loss_local = float((-(t * np.log(pc + eps) + (1 - t) * np.log(1 - pc + eps))).sum() * inv)
            dp_ = -(t / (pc + eps) - (1 - t) / (1 - pc + eps)) * inv
            self.dz2 = (dp_ * self.p * (1 - self.p)).astype(np.float32)
            g[:self.off_w] = (self.dz2.T @ self.h).ravel()             # dW'^T, item-major
        elif mask == 16:                                               # GRADS_C
            dz1 = ((self.dz2 @ o.W_.T) * self.h * (1 - self.h)).astype(np.float32)
            self.dz1.numpy()[:] = dz1
            a = self.off_w
            g[a:a + I * K] = (self.x.T @ dz1).ravel()
            g[a + I * K:a + I * K + K] = dz1.sum(0)
            g[a + I * K + K:self.off_v] = self.dz2.sum(0)
        elif mask in (128, 256, 512):                                  # the three update launches
            c = np.float32(self.reg / self.gbatch)
            t = lambda j: 5 * (self.step - 1) + j + 1
            a = self.off_w
            if mask == 128:
                gv = g[self.off_v:].reshape(self.U, K) + c * o.V
                self.adam_update(o.V, o.m[2], o.v[2], gv.astype(np.float32), o.lr, t(2))
            elif mask == 256:
                gw_ = g[:self.off_w].reshape(I, K).T + c * o.W_
                self.reg_w_ = float(c * 0.5 * (o.W_.astype(np.float64) ** 2).sum())
                self.adam_update(o.W_, o.m[1], o.v[1], np.ascontiguousarray(gw_, np.float32), o.lr, t(1))
            else:
                gw = g[a:a + I * K].reshape(I, K) + c * o.W
                self.adam_update(o.W, o.m[0], o.v[0], gw.astype(np.float32), o.lr, t(0))
                self.adam_update(o.b, o.m[3], o.v[3], g[a + I * K:a + I * K + K].copy(), o.lr, t(3))
                self.adam_update(o.b_, o.m[4], o.v[4], g[a + I * K + K:self.off_v].copy(), o.lr, t(4))
        else:
            raise AssertionError(mask)

    def add_user_rows(self, uids_all, rows_all):
        gv = self.grads.numpy()[self.off_v:].reshape(self.U, self.K)
        np.add.at(gv, uids_all.numpy(), rows_all.numpy())


def _schedule_worker(rank, world, port, out):
    from drecpy_b200.parallel import data_parallel_step
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dp = DataParallel(dist)
    U, I, K, Bg, csr, w, uids, keep = _problem()
    lo, hi = dp.shard(Bg)
    ph = OraclePhases(w, csr, uids[lo:hi], keep[lo:hi], Bg)
    rows_all = torch.zeros((Bg, K), dtype=torch.float32)
    uids_all = torch.zeros(Bg, dtype=torch.int32)
    data_parallel_step(dp, ph, ph.label_count, ph.grads, ph.off_w, ph.off_v, ph.dz1,
                       torch.from_numpy(uids[lo:hi].astype(np.int32)), rows_all, uids_all, ph.add_user_rows)
    assert ph.log == [1, 2, 8, 16, 128, 256, 512]
    assert np.array_equal(uids_all.numpy(), uids)                     # gathered in rank order == the global batch
    o = ph.o
    np.savez(out + f'.{rank}.npz', W=o.W, W_=o.W_, V=o.V, b=o.b, b_=o.b_)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_schedule_equals_single_process_step(tmp_path):
    out = str(tmp_path / 'sched')
    mp.spawn(_schedule_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    U, I, K, Bg, csr, w, uids, keep = _problem()
    o = CDAEOracle(*w, csr, learning_rate=1e-3)
    o.step(uids, keep, 1e-3)                                          # one process, the whole global batch
    r0, r1 = np.load(out + '.0.npz'), np.load(out + '.1.npz')
    for name in ('W', 'W_', 'V', 'b', 'b_'):
        assert np.array_equal(r0[name], r1[name]), name               # replicas stay identical
        want = getattr(o, name)
        assert np.abs(r0[name] - want).max() <= 2e-4 * np.abs(want).max() + 1e-7, name
