"""The drop-in claim, tested from the caller's side (-m gpu): the native models are fitted from objects shaped like
the reference's InteractionDataset (not this package's InteractionData), and the reference's per-user evaluation loop
-- restated here from DRecPy/Evaluation/Processes/ranking_evaluation.py:107-116 (4-thread pool, random.Random(seed+idx)
per user) and :163-246 (candidates, model.rank, relevancies, metrics) -- drives model.rank concurrently.  It touches
only what the reference touches: model.rank, model.interaction_dataset.select(...), model.interaction_threshold,
model.n_items."""
import random
import threading
from multiprocessing.pool import ThreadPool

import numpy as np
import pandas as pd
import pytest

import drecpy_b200 as drb
from oracle.ranking import hit_ratio, ndcg

pytestmark = pytest.mark.gpu


class FrameDataset:
    """Looks like DRecPy's MemoryInteractionDataset to fit(): a `_df` with user / item / interaction columns
    (mem_dataset.py:37-60), assign_internal_ids(), select(), unique(), values_list(), __len__."""

    def __init__(self, df):
        self._df = df.reset_index(drop=True)
        self.has_internal_ids = False

    def assign_internal_ids(self):
        self.has_internal_ids = True

    def __len__(self):
        return len(self._df)

    def select(self, query):                       # 'user == 5, interaction >= 0.001' (mem_dataset.py:62-99)
        df = self._df
        for part in query.split(','):
            col, op, val = part.split()
            val = float(val)
            df = df[{'==': df[col] == val, '>=': df[col] >= val, '<': df[col] < val}[op]]
        return FrameDataset(df)

    def unique(self, col):
        return FrameDataset(self._df.drop_duplicates(col))

    def values_list(self, cols=None, to_list=False):
        cols = [cols] if isinstance(cols, str) else list(cols or self._df.columns)
        if to_list:
            rows = self._df[cols].values.tolist()
            return [r[0] for r in rows] if len(cols) == 1 else rows
        return self._df[cols].to_dict('records')

    def select_one(self, query, cols, to_list=True):
        sub = self.select(query)
        return None if len(sub) == 0 else sub._df[cols[0]].iloc[0]


class ListDataset:
    """The other branch of InteractionData.from_dataset: no `_df`, only the generic InteractionDataset surface
    (values_list / select / unique), as the reference's DatabaseInteractionDataset offers."""

    def __init__(self, rows):
        self.rows = list(rows)

    def assign_internal_ids(self):
        pass

    def __len__(self):
        return len(self.rows)

    def values_list(self, cols=None, to_list=False):
        idx = [('user', 'item', 'interaction').index(c) for c in cols]
        return [[r[j] for j in idx] for r in self.rows] if to_list else [{c: r[j] for c, j in zip(cols, idx)} for r in self.rows]

    def select(self, query):
        return FrameDataset(pd.DataFrame(self.rows, columns=['user', 'item', 'interaction'])).select(query)

    def unique(self, col):
        return FrameDataset(pd.DataFrame(self.rows, columns=['user', 'item', 'interaction'])).unique(col)


def _split(seed=12, U=150, I=260, nnz=7000):
    u, i, v = drb.synthetic_interactions(U, I, nnz, seed=seed)
    rng = np.random.default_rng(1)
    test_mask = np.zeros(len(u), bool)
    for usr in np.unique(u):
        idx = np.flatnonzero(u == usr)
        if len(idx) > 3: test_mask[rng.choice(idx)] = True
    mk = lambda sel: pd.DataFrame({'user': u[sel], 'item': i[sel], 'interaction': v[sel]})
    return mk(~test_mask), mk(test_mask)


def reference_user_task(model, user, ds_test, thr, n_pos, n_neg, metrics, novelty, sums, k, rng, lock, record):
    """ranking_evaluation.py:163-246 for one user, against a dataset object exposing select / values_list."""
    user_ds = ds_test.select(f'user == {user}')
    pos_ds = user_ds.select(f'interaction >= {thr}')
    if len(pos_ds) < n_pos:
        return
    chosen = rng.sample(pos_ds.values_list(['item', 'interaction']), n_pos)
    positives = [p['item'] for p in chosen]
    neg_ds = user_ds.select(f'interaction < {thr}')
    negatives = rng.sample(neg_ds.values_list(['item'], to_list=True), min(n_neg, len(neg_ds)))
    if len(negatives) < n_neg:
        train_pos = model.interaction_dataset.select(f'user == {user}, interaction >= {thr}')
        blacklist = set(train_pos.unique('item').values_list('item', to_list=True))
        blacklist |= set(pos_ds.unique('item').values_list('item', to_list=True))
        if model.n_items - len(blacklist) < n_neg:
            return
        while len(negatives) < n_neg:
            new_item = rng.randint(0, model.n_items - 1)
            if new_item not in blacklist and new_item not in negatives:
                negatives.append(new_item)
    all_items = positives + negatives
    rng.shuffle(all_items)
    recommendations = [item for _, item in model.rank(user, all_items, novelty=novelty, skip_invalid_items=True)]
    relevancies = {item: (user_ds.select_one(f'item == {item}', ['interaction'], to_list=True) or 0)
                   for item in all_items}
    with lock:
        record[user] = (all_items, recommendations)
        for name, fn in metrics:
            try:
                if name == 'HitRatio':
                    sums[name][0] += fn(recommendations, k=k, relevant_recommendations=positives)
                else:
                    sums[name][0] += fn(recommendations, k=k, relevancies=relevancies)
                sums[name][1] += 1
            except Exception:
                pass


@pytest.mark.parametrize('model_kind', ['cdae', 'dmf'])
def test_reference_shaped_dataset_and_threaded_rank_loop(model_kind):
    train_df, test_df = _split()
    frame = FrameDataset(train_df)
    rows = ListDataset(train_df[['user', 'item', 'interaction']].values.tolist())
    native = drb.InteractionData(train_df['user'].values, train_df['item'].values, train_df['interaction'].values)

    def make():
        if model_kind == 'cdae':
            return drb.CDAE(hidden_factors=20, seed=10, verbose=False), dict(epochs=12, batch_size=32)
        return drb.DMF(user_factors=[32, 16], item_factors=[32, 16], seed=10, verbose=False), \
            dict(epochs=12, batch_size=64, reg_rate=1e-4)
    models = []
    for ds in (frame, rows, native):
        m, kw = make()
        m.fit(ds, **kw)
        models.append(m)
    # the three ways of handing over the same rows give the same fitted model (same seed, same sampler stream)
    # (the scatter kernels add with floating-point atomics, so equal up to summation order, not bit for bit)
    ref = models[2]._params.cpu().numpy()
    for m in models[:2]:
        assert np.allclose(m._params.cpu().numpy(), ref, rtol=1e-4, atol=1e-6)
    model = models[0]
    assert model.interaction_dataset is frame          # what ranking_evaluation.py:196 reads

    # the reference's driver loop (ranking_evaluation.py:107-116): 4 threads, Random(seed) with seed += 1 per user
    ds_test = FrameDataset(test_df)
    users = ds_test.unique('user').values_list(['user'], to_list=True)
    thr, seed, k = model.interaction_threshold, 10, 10
    metrics = [('HitRatio', hit_ratio), ('NDCG', ndcg)]
    sums = {name: [0, 0] for name, _ in metrics}
    record, lock = {}, threading.Lock()
    pool = ThreadPool(processes=4)
    for idx, user in enumerate(users):
        pool.apply_async(reference_user_task, (model, user, ds_test, thr, 1, 50, metrics, True, sums, k,
                                               random.Random(seed + idx), lock, record))
    pool.close()
    pool.join()
    assert len(record) == len(users)
    threaded = {f'{n}@{k}': round(sums[n][0] / sums[n][1], 4) for n in sums}

    # the same loop, one thread: concurrent rank() calls must not have disturbed each other
    for idx, user in enumerate(users[:40]):
        single = {}
        reference_user_task(model, user, ds_test, thr, 1, 50, metrics, True, {n: [0, 0] for n in sums}, k,
                            random.Random(seed + idx), threading.Lock(), single)
        assert single[user] == record[user]

    # and the batched native evaluator reports the same numbers as the reference-shaped loop
    test_native = drb.InteractionData(test_df['user'].values, test_df['item'].values, test_df['interaction'].values)
    got = drb.ranking_evaluation(model, test_native, k=k, n_pos_interactions=1, n_neg_interactions=50,
                                 generate_negative_pairs=True, novelty=True, seed=seed,
                                 metrics=[drb.HitRatio(), drb.NDCG()], verbose=False)
    assert got == threaded, (got, threaded)
    got_frame = drb.ranking_evaluation(model, ds_test, k=k, n_pos_interactions=1, n_neg_interactions=50,
                                       generate_negative_pairs=True, novelty=True, seed=seed,
                                       metrics=[drb.HitRatio(), drb.NDCG()], verbose=False)
    assert got_frame == threaded
