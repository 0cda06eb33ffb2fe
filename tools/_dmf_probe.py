import sys, os, time, json, gc
sys.path.insert(0, '.')
import numpy as np, torch
import drecpy_b200 as drb
def run(tag, nogc):
    u, i, v = drb.synthetic_interactions(6040, 3706, 1_000_000, seed=10)
    ds = drb.InteractionData(u, i, v)
    m = drb.DMF(user_factors=[64, 32], item_factors=[64, 32], seed=10, verbose=False)
    B = 256
    m.fit(ds, epochs=0, batch_size=B, learning_rate=1e-3, neg_ratio=5, reg_rate=1e-4)
    if nogc: gc.disable()
    per = []
    t0 = time.perf_counter()
    for s in range(4000):
        t1 = time.perf_counter()
        m._step += 1; m._train_step(B, 1e-4, want_loss=False, prefetch=True)
        per.append(time.perf_counter() - t1)
    m.synchronize()
    tot = time.perf_counter() - t0
    per = np.array(per) * 1e6
    big = [(int(k), round(float(per[k]), 0)) for k in np.flatnonzero(per > 1000)]
    print(tag, 'nogc' if nogc else 'gc', 'avg_us', round(tot / 4000 * 1e6, 1), 'median', round(float(np.median(per)), 1), 'calls > 1 ms:', big)
    gc.enable()
run('GRAPH', False)
run('GRAPH', True)
os.environ['DRB_GRAPH'] = '0'
run('DIRECT', False)
