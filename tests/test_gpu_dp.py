"""Data-parallel CDAE step on 2 GPUs (NCCL) against the single-GPU step with the same global batch (P9)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _weights(U, I, K):
    rng = np.random.default_rng(5)

    def glorot(shape, fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    return {'W': glorot((I, K), I, K), 'W_': glorot((K, I), K, I), 'V': glorot((U, K), U, K),
            'b': glorot((K,), K, K), 'b_': glorot((I,), I, I)}


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import drecpy_b200 as drb
    from drecpy_b200.parallel import DataParallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    U, I, K, B = 500, 900, 64, 96
    u, i, v = drb.synthetic_interactions(U, I, 30000, seed=4)
    ds = drb.InteractionData(u, i, v)
    w = _weights(U, I, K)
    m = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='philox', device=f'cuda:{rank}')
    m.fit(ds, epochs=0, batch_size=B, init_weights=w, data_parallel=DataParallel(dist))
    losses = []
    for s in range(1, 7):
        m._step = s
        losses.append(m._train_step(B, 1e-3, want_loss=True))
    if rank == 0:
        ref = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='philox', device='cuda:0')
        ref.fit(ds, epochs=0, batch_size=B * world, init_weights=w)
        ref_losses = []
        for s in range(1, 7):
            ref._step = s
            ref_losses.append(ref._train_step(B * world, 1e-3, want_loss=True))
        np.savez(out, losses=losses, ref_losses=ref_losses, p=m._params.cpu().numpy(), p_ref=ref._params.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_step_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'dp.npz')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert np.allclose(r['losses'], r['ref_losses'], rtol=1e-5), (r['losses'], r['ref_losses'])
    scale = np.abs(r['p_ref']).max()
    assert np.abs(r['p'] - r['p_ref']).max() < 2e-4 * scale


def _worker_items(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import drecpy_b200 as drb
    from drecpy_b200.parallel import DataParallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    U, I, K, B = 501, 903, 64, 96          # odd sizes: uneven shards
    u, i, v = drb.synthetic_interactions(U, I, 30000, seed=4)
    ds = drb.InteractionData(u, i, v)
    w = _weights(U, I, K)
    m = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='philox', device=f'cuda:{rank}')
    m.fit(ds, epochs=0, batch_size=B, init_weights=w, data_parallel=DataParallel(dist), parallel_mode='items')
    assert m.W.shape[0] < I and m.V.shape[0] < U          # weights really are sharded
    losses = []
    for s in range(1, 7):
        m._step = s
        losses.append(m._train_step(B, 1e-3, want_loss=True))
    full = m.gather_full_weights()
    if rank == 0:
        ref = drb.CDAE(hidden_factors=K, seed=10, verbose=False, rng_mode='philox', device='cuda:0')
        ref.fit(ds, epochs=0, batch_size=B * world, init_weights=w)
        ref_losses = []
        for s in range(1, 7):
            ref._step = s
            ref_losses.append(ref._train_step(B * world, 1e-3, want_loss=True))
        np.savez(out, losses=losses, ref_losses=ref_losses,
                 **{k: t.cpu().numpy() for k, t in full.items()},
                 **{'ref_' + k: getattr(ref, k).cpu().numpy() for k in ('W', 'W_', 'V', 'b', 'b_')})
    dist.barrier()
    dist.destroy_process_group()


def test_item_sharded_step_matches_single_gpu(tmp_path):
    """W, W', b' sharded by item range and V by user range over 2 GPUs; only batch x hidden activations travel."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'items.npz')
    mp.spawn(_worker_items, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert np.allclose(r['losses'], r['ref_losses'], rtol=1e-5), (r['losses'], r['ref_losses'])
    for k in ('W', 'W_', 'V', 'b', 'b_'):
        scale = np.abs(r['ref_' + k]).max()
        assert r[k].shape == r['ref_' + k].shape
        assert np.abs(r[k] - r['ref_' + k]).max() < 2e-4 * scale, k


def _worker_dmf(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import drecpy_b200 as drb
    from drecpy_b200.parallel import DataParallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    U, I, B = 400, 600, 64
    u, i, v = drb.synthetic_interactions(U, I, 20000, seed=5)
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    rng = np.random.default_rng(2)

    def tower(in_dim, factors):
        res = []
        for f in factors:
            lim = np.sqrt(6.0 / (in_dim + f))
            res.append((rng.uniform(-lim, lim, (in_dim, f)).astype(np.float32), np.zeros(f, np.float32)))
            in_dim = f
        return res
    w = {'user_nn': tower(I, [64, 32]), 'item_nn': tower(U, [64, 32])}
    m = drb.DMF(seed=10, verbose=False, device=f'cuda:{rank}')
    m.fit(ds, epochs=0, batch_size=B, reg_rate=1e-4, init_weights=w, data_parallel=DataParallel(dist))
    # every rank takes its slice of the SAME global sample stream
    gs = drb.PointSampler(ds, 5, 1e-3, 10)
    losses = []
    dev = torch.device(f'cuda:{rank}')
    loss = torch.zeros(2, device=dev)
    for s in range(6):
        uu, ii, vv = gs.sample_arrays(B * world)
        lab = m.labels_from_values(vv)
        sl = slice(rank * B, (rank + 1) * B)
        m.step_device(torch.as_tensor(uu[sl].copy(), device=dev), torch.as_tensor(ii[sl].copy(), device=dev),
                      torch.as_tensor(lab[sl].copy(), device=dev), 1e-4, loss)
        losses.append(m.global_loss(loss))
    if rank == 0:
        ref = drb.DMF(seed=10, verbose=False, device='cuda:0')
        ref.fit(ds, epochs=0, batch_size=B * world, reg_rate=1e-4, init_weights=w)
        gs = drb.PointSampler(ds, 5, 1e-3, 10)
        ref_losses = []
        rl = torch.zeros(2, device=dev)
        for s in range(6):
            uu, ii, vv = gs.sample_arrays(B * world)
            ref.step_device(torch.as_tensor(uu.copy(), device=dev), torch.as_tensor(ii.copy(), device=dev),
                            torch.as_tensor(ref.labels_from_values(vv), device=dev), 1e-4, rl)
            ref_losses.append(float(rl[0]))
        np.savez(out, losses=losses, ref_losses=ref_losses, p=m._params.cpu().numpy(), p_ref=ref._params.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_dmf_two_gpu_step_matches_single_gpu(tmp_path):
    """DMF data parallel: 2 ranks x 64 pairs == one rank x 128 pairs (same pairs, global loss mean, one all-reduce)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'dmf_dp.npz')
    mp.spawn(_worker_dmf, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert np.allclose(r['losses'], r['ref_losses'], rtol=1e-5), (r['losses'], r['ref_losses'])
    scale = np.abs(r['p_ref']).max()
    assert np.abs(r['p'] - r['p_ref']).max() < 2e-4 * scale


def _worker_items_sampled(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import drecpy_b200 as drb
    from drecpy_b200.parallel import DataParallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    U, I, K, B = 501, 904, 64, 96
    u, i, v = drb.synthetic_interactions(U, I, 30000, seed=4)
    ds = drb.InteractionData(u, i, v)
    w = _weights(U, I, K)
    kw = dict(hidden_factors=K, seed=10, verbose=False, rng_mode='philox', output='sampled', neg_per_group=24)
    m = drb.CDAE(device=f'cuda:{rank}', **kw)
    m.fit(ds, epochs=0, batch_size=B, init_weights=w, data_parallel=DataParallel(dist), parallel_mode='items')
    losses = []
    for s in range(1, 7):
        m._step = s
        losses.append(m._train_step(B, 1e-3, want_loss=True))
    full = m.gather_full_weights()
    if rank == 0:
        ref = drb.CDAE(device='cuda:0', neg_groups=world, **kw)       # one negative-sampling group per shard
        ref.fit(ds, epochs=0, batch_size=B * world, init_weights=w)
        ref_losses = []
        for s in range(1, 7):
            ref._step = s
            ref_losses.append(ref._train_step(B * world, 1e-3, want_loss=True))
        np.savez(out, losses=losses, ref_losses=ref_losses,
                 **{k: t.cpu().numpy() for k, t in full.items()},
                 **{'ref_' + k: getattr(ref, k).cpu().numpy() for k in ('W', 'W_', 'V', 'b', 'b_')})
    dist.barrier()
    dist.destroy_process_group()


def test_item_sharded_sampled_output_step_matches_single_gpu(tmp_path):
    """configs[4]'s mode at toy size: item-sharded weights + sampled outputs on 2 GPUs == one GPU with two
    negative-sampling groups (same draws, same positives), only batch x hidden activations travel."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'items_sampled.npz')
    mp.spawn(_worker_items_sampled, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert np.allclose(r['losses'], r['ref_losses'], rtol=1e-5), (r['losses'], r['ref_losses'])
    for k in ('W', 'W_', 'V', 'b', 'b_'):
        scale = np.abs(r['ref_' + k]).max()
        assert np.abs(r[k] - r['ref_' + k]).max() < 2e-4 * scale, k
