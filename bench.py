#!/usr/bin/env python
"""bench.py -- training samples/s of the CDAE step on the synthetic ml-20m shape (BASELINE.json configs[2]).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference-semantics CPU path, host cores)

One "step" = one optimizer step of the hot path on one batch of 4096 sampled users per GPU (weak scaling).
Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA-event timed; `e2e` = the public per-step
path (host sampler -> pinned H2D -> step -> D2H loss); `roofline` = the dominant kernel family timed live with
CUDA events inside libdrb (separate pass); `cpu_baseline` = the oracle port timed on the host cores (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(name='cdae_ml20m_shape', n_users=138493, n_items=26744, nnz=20_000_000, hidden=200, batch=4096,
          q=0.2, lr=1e-3, reg=1e-3, neg_ratio=5, seed=10, zipf_a=1.0)
SMALL = dict(name='cdae_small_debug', n_users=6040, n_items=3706, nnz=1_000_000, hidden=200, batch=1024,
             q=0.2, lr=1e-3, reg=1e-3, neg_ratio=5, seed=10, zipf_a=1.0)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf=d['bf16_tflops_sustained'], tf_burst=d['bf16_tflops'], src='measured')
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src='fallback')


def build_roofline(kernels, hidden, n_items, batch, n_params, peaks, traffic_table=None):
    """The `roofline` object of the bench line from the per-kernel CUDA-event times (ms per step).

    Dominant kernels: the three output-layer GEMMs (forward + fused loss epilogue, dW'^T, dh), tensor bound.
    `achieved` = ALGORITHMIC flops (6*K*I per sampled user) / their time; `peak` = the measured sustained
    bf16 rate.  Every fp32-accurate product is issued as three TF32 MMAs and TF32 runs at half the bf16 rate, so this
    design can reach at most peak / 6 of algorithmic flops (`cap_3xtf32`); `frac_of_cap` says how close the kernels are
    to that."""
    gemm_ms = sum(v for k, v in kernels.items() if k.startswith('k_sgemm') or k.startswith('k_umma'))
    flops = 6.0 * hidden * n_items * batch
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    tc_path = any(k.startswith('k_umma') for k in kernels)
    traffic = None
    if tc_path and traffic_table:
        traffic = sum(traffic_table[k] for k in ('k_umma_cdae_loss', 'k_umma_gemm_mn', 'k_umma_gemm_kk'))
    adam_ms = kernels.get('k_adam', float('nan'))
    adam_gbs = 28.0 * n_params / (adam_ms * 1e-3) / 1e9
    cap = peaks['tf'] / 6.0 if tc_path else None
    return {'kernel': ('k_umma_cdae_loss + k_umma_gemm x2 (tcgen05 3xTF32: output layer fwd + fused loss epilogue, '
                       'dW\'^T and dh)') if tc_path else 'k_sgemm x3 (fp32 FFMA path)',
            'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tf'], 'unit': 'TFLOP/s',
            'frac': (achieved / peaks['tf']) if achieved else None, 'traffic': traffic,
            'note': ('achieved counts the algorithmic fp32 flops (6*K*I per sampled user); every product is issued '
                     'as 3 TF32 MMAs (half the bf16 rate each), so the issued-MMA rate is 3x achieved against a '
                     'TF32 peak of half the bf16 peak: at most peak/6 of algorithmic flops') if tc_path
            else 'CUDA-core path',
            'achieved_issued_tf32': 3 * achieved if (achieved and tc_path) else None,
            'cap_3xtf32': cap, 'frac_of_cap': (achieved / cap) if (achieved and cap) else None,
            'peak_source': f"{peaks['src']} bf16 sustained",
            'share_of_step': gemm_ms / sum(kernels.values()) if kernels else None,
            'secondary': {'k_adam': {'bound': 'hbm', 'achieved': adam_gbs, 'peak': peaks['hbm'], 'unit': 'GB/s',
                                     'frac': adam_gbs / peaks['hbm']}}}


def make_data(cfg):
    import drecpy_b200 as drb
    t = time.time()
    u, i, v = drb.synthetic_interactions(cfg['n_users'], cfg['n_items'], cfg['nnz'], seed=cfg['seed'],
                                         zipf_a=cfg['zipf_a'])
    ds = drb.InteractionData(u, i, v)
    ds.assign_internal_ids()
    ds.csr(1e-3)
    return ds, time.time() - t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].startswith('Active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def oracle_model(cfg, ds):
    from oracle.cdae import CDAEOracle
    rng = np.random.default_rng(1)
    U, I, K = cfg['n_users'], cfg['n_items'], cfg['hidden']

    def glorot(shape, fi, fo):
        lim = np.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    return CDAEOracle(glorot((I, K), I, K), glorot((K, I), K, I), glorot((U, K), U, K), glorot((K,), K, K),
                      glorot((I,), I, I), ds.csr(), interaction_threshold=1e-3, corruption_level=cfg['q'],
                      learning_rate=cfg['lr'])


def time_oracle(cfg, ds, steps, warmup, budget_s=150.0):
    """Reference-semantics CPU path (oracle port: batch-mean labels, per-variable Adam counter, dense Adam + L2),
    numpy / BLAS on all host cores.  The live reference sampler is replaced by the oracle-equivalent native one and
    the corruption mask by a numpy Bernoulli draw (both cheaper than the reference's Python loops)."""
    import drecpy_b200 as drb
    o = oracle_model(cfg, ds)
    sampler = drb.PointSampler(ds, cfg['neg_ratio'], 1e-3, cfg['seed'])
    rng = np.random.default_rng(0)
    B = cfg['batch']

    def one(b):
        uids = sampler.sample_arrays(b)[0]
        keep = rng.random((b, cfg['n_items'])) >= cfg['q']
        return float(o.step(uids, keep, cfg['reg']))
    t0 = time.time()
    one(min(B, 512))                                   # calibration: per-user cost + fixed dense-Adam cost
    t_cal = time.time() - t0
    est_full = t_cal * max(1.0, B / 512 * 0.6)
    b_ref = B
    while b_ref > 256 and est_full * (b_ref / B) * (steps + warmup) > budget_s:
        b_ref //= 2
    for _ in range(warmup):
        one(b_ref)
    t0 = time.time()
    for _ in range(steps):
        loss = one(b_ref)
    dt = time.time() - t0
    return dict(samples_per_s=b_ref * steps / dt, batch=b_ref, seconds=dt, loss=loss, steps=steps)


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ds, _ = make_data(cfg)
    steps = max(1, args.steps)
    r = time_oracle(cfg, ds, steps, args.warmup)
    cores = os.cpu_count()
    sample = f"{steps} timed oracle steps of {r['batch']} sampled users on the full {cfg['name']} shape"
    line = {'impl': 'reference', 'metric': 'cdae_training_samples_per_sec', 'value': r['samples_per_s'],
            'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * r['seconds'] / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(cfg, args.gpus),
            'cpu_baseline': {'value': r['samples_per_s'], 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': r['samples_per_s'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, n_gpus, parallelism=None):
    parallelism = parallelism or f'dp{n_gpus}'
    origin = 'BASELINE.json configs[2]' if cfg['name'] == C3['name'] else 'debug shape, not a BASELINE.json config'
    return {'workload': f"{cfg['name']}: CDAE hidden_factors={cfg['hidden']} bce q={cfg['q']} on synthetic "
                        f"{cfg['n_users']}x{cfg['n_items']} / {cfg['nnz']} interactions ({origin})",
            'batch_per_gpu': cfg['batch'], 'global_batch': cfg['batch'] * n_gpus, 'neg_ratio': cfg['neg_ratio'],
            'label_mode': 'batch_mean', 'adam': 'dense, per-variable step counter', 'mask_rng': 'philox (device)',
            'item_popularity': f"zipf a={cfg['zipf_a']}", 'parallelism': parallelism,
            'l2': ('per-step working set (params + Adam state + dz ~ 2.7 GB) exceeds the 126 MB L2; no flush needed'
                   if cfg['name'] == C3['name'] else 'debug shape: the working set is L2 resident')}


# ------------------------------------------------------------------------------------------ GPU arm
def run_native(args, cfg):
    import torch
    import drecpy_b200 as drb
    from drecpy_b200 import _lib
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    dev = torch.device(f'cuda:{local}')
    ds, t_data = make_data(cfg)
    B, K, W = cfg['batch'], args.steps, args.warmup
    m = drb.CDAE(hidden_factors=cfg['hidden'], corruption_level=cfg['q'], loss='bce', seed=cfg['seed'], verbose=False,
                 rng_mode='philox', device=str(dev))
    from drecpy_b200.parallel import DataParallel
    m.fit(ds, epochs=0, batch_size=B, learning_rate=cfg['lr'], neg_ratio=cfg['neg_ratio'], reg_rate=cfg['reg'],
          sampler=drb.PointSampler(ds, cfg['neg_ratio'], 1e-3, cfg['seed'] + rank),
          data_parallel=DataParallel(dist), dp_sampler='independent', parallel_mode=args.parallel)
    items_mode = args.parallel == 'items' and world > 1
    Bs = B * world if items_mode else B          # item-sharded: every rank steps the whole global batch

    # ---- device-resident inputs for the `value` leg
    lib = _lib.load()
    pos_indptr = np.ascontiguousarray(ds.csr(1e-3)[0])
    batches = []
    pos_indptr = m._h_indptr                      # the (possibly item-sharded) CSR the model gathers from
    if items_mode:                                # all ranks must step the same users
        m._sampler = drb.PointSampler(ds, cfg['neg_ratio'], 1e-3, cfg['seed'])
    for _ in range(K + W):
        u = m._sampler.sample_arrays(Bs)[0]
        off = np.zeros(Bs + 1, np.int32)
        _lib.check(lib.drb_batch_offsets(_lib.np_ptr(u), Bs, _lib.np_ptr(pos_indptr), _lib.np_ptr(off)))
        batches.append((torch.from_numpy(u.copy()).to(dev), torch.from_numpy(off).to(dev)))
    loss_dev = torch.zeros(2, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        m.step_device(batches[s][0], batches[s][1], None, cfg['reg'], loss_dev)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = m.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        m.step_device(batches[s][0], batches[s][1], None, cfg['reg'], loss_dev)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clock_info = clocks.stop()
    launches = m.launch_count() - launches0
    loss_value = m.global_loss(loss_dev)                        # collective when parallel: every rank calls it
    t = torch.tensor([ms_total], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- e2e leg: the per-step body of fit(): host sampler -> pinned staging -> H2D -> step -> D2H loss
    for _ in range(2):
        m._step += 1
        m._train_step(B, cfg['reg'], want_loss=True, prefetch=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        m._step += 1
        loss_e2e = m._train_step(B, cfg['reg'], want_loss=True, prefetch=True)   # exactly what fit() runs per epoch
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())

    # ---- roofline leg: per-kernel CUDA-event timing inside libdrb (separate pass; every rank runs it because the
    # step contains collectives, rank 0 reports)
    P = 5
    _lib.check(lib.drb_ctx_profile_enable(m._ctx, 1))
    for s in range(P):
        bt = batches[s % len(batches)]
        m.step_device(bt[0], bt[1], None, cfg['reg'], loss_dev)
    prof = _lib.profile_read(m._ctx)
    _lib.check(lib.drb_ctx_profile_enable(m._ctx, 0))
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    kernels = {k: round(v[0] / P, 4) for k, v in prof.items()}
    peaks = load_peaks()
    I, Hd, U = cfg['n_items'], cfg['hidden'], cfg['n_users']
    traffic_table = None
    tpath = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.exists(tpath) and cfg['name'] == C3['name']:
        traffic_table = json.load(open(tpath))['dram_bytes_per_launch']   # ncu --set full capture of the same command
    n_params = int(m._L.total)                                   # parameters this rank updates (its shard when item-sharded)
    launches_total = int(launches)
    roofline = build_roofline(kernels, Hd, I, B, n_params, peaks, traffic_table)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_oracle(cfg, ds, 2, 1, budget_s=60.0)
        cpu = {'value': r['samples_per_s'], 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': f"2 timed oracle steps of {r['batch']} sampled users on the full shape (numpy/BLAS, all cores)"}

    extras = None
    if world == 1 and not args.no_extras:
        # the other two legs of BASELINE.json's metric, measured after the headline (not part of `value`):
        # DMF training samples/s at configs[1] and ranked users/s at configs[3] (tools/bench_extra.py)
        try:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import bench_extra
            del m, batches
            torch.cuda.empty_cache()
            extras = bench_extra.dmf_c2(steps=100)
            extras.update(bench_extra.eval_c4(K=cfg['hidden'], arrays=(ds.user, ds.item, ds.interaction)))
        except Exception as exc:          # never lose the headline line because of a secondary measurement
            extras = {'error': repr(exc)}

    value = B * world * K / (ms_total * 1e-3)
    line = {'metric': 'cdae_training_samples_per_sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(cfg, world, f'item-sharded x{world}' if items_mode else f'dp{world}'),
            'clocks': clock_info, 'gpu_launches': launches_total,
            'e2e': {'value': B * world * K / t_e2e, 'unit': 'samples/s', 'h2d_bytes_per_step': 4 * B + 4 * (B + 1),
                    'd2h_bytes_per_step': 4, 'ms_per_step': 1e3 * t_e2e / K},
            'roofline': roofline, 'cpu_baseline': cpu, 'kernels_ms_per_step': kernels,
            'loss_last': loss_value, 'loss_last_e2e': loss_e2e, 'data_gen_s': round(t_data, 1), 'extras': extras}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='c3', choices=['c3', 'small'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--parallel', default='data', choices=['data', 'items'],
                    help='N>1: data = replicated weights + gradient all-reduce; items = item-sharded weights')
    args = ap.parse_args()
    cfg = dict(C3 if args.workload == 'c3' else SMALL)
    if args.impl == 'reference':
        run_reference(args, cfg)
    else:
        run_native(args, cfg)


if __name__ == '__main__':
    main()
