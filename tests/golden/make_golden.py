"""Generate the golden vectors under tests/golden/ by running the LIVE reference (/root/reference).

Run in the build container only:   python tests/golden/make_golden.py
The reference is imported through tests/golden/ref_shim.py (numpy.float alias + tensorflow/matplotlib stubs);
none of the code exercised here touches TensorFlow.  Outputs (committed):
  sampler_<name>.npz   dataset rows + PointSampler(ds, neg_ratio, thr, seed).sample(n) triples
                       (DRecPy/Sampler/point_sampler.py:44-96) + the scipy interaction matrix the reference
                       builds (DRecPy/Dataset/mem_dataset.py:480-498) + internal ids (:309-330)
  ranking.json         candidate lists handed to model.rank and the final metric dicts of the live
                       ranking_evaluation (DRecPy/Evaluation/Processes/ranking_evaluation.py:19-246) on the
                       reference's own fixture (tests/Evaluation/Processes/test_ranking_evaluation.py:12-19)
  splits.json          train / test row ids of the live leave_k_out (DRecPy/Evaluation/Splits/leave_k_out.py:14-135):
                       fixed k, ratio k, min_user_interactions, k larger than some users' histories,
                       last_timestamps
  ref_kat_check.json   record that the oracle protocol driven by the live UserKNN reproduces the reference's
                       own expected metric dicts (test_ranking_evaluation.py:30-60)
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

ref_shim.install()

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
from DRecPy.Dataset import InteractionDataset  # noqa: E402
from DRecPy.Sampler import PointSampler  # noqa: E402
from DRecPy.Evaluation.Processes import ranking_evaluation  # noqa: E402
from DRecPy.Evaluation.Splits import leave_k_out  # noqa: E402
from DRecPy.Evaluation.Metrics import HitRatio, NDCG, Precision, Recall  # noqa: E402


def synth(U, I, nnz, seed, zero_frac=0.0, dup=0):
    rng = np.random.default_rng(seed)
    pairs = rng.choice(U * I, nnz, replace=False)
    user = pairs // I + 1
    item = pairs % I + 1
    val = rng.integers(1, 6, nnz)
    if zero_frac > 0:
        val[rng.random(nnz) < zero_frac] = 0
    if dup:
        pick = rng.choice(nnz, dup, replace=False)
        user = np.concatenate([user, user[pick]])
        item = np.concatenate([item, item[pick]])
        val = np.concatenate([val, rng.integers(1, 6, dup)])
    # make sure every user / item id appears (nominal shape) -- not required, shape is read back from the data
    return user.astype(np.int64), item.astype(np.int64), val.astype(np.int64)


def sampler_golden(name, U, I, nnz, data_seed, seeds, n, neg_ratio=5, thr=0.001, zero_frac=0.0, dup=0,
                   float_vals=False):
    user, item, val = synth(U, I, nnz, data_seed, zero_frac, dup)
    if float_vals:
        val = val.astype(np.float64) + 0.5
    df = pd.DataFrame({'user': user, 'item': item, 'interaction': val})
    ds = InteractionDataset.read_df(df, verbose=False)
    ds.assign_internal_ids()
    out = {'user': user, 'item': item, 'interaction': val,
           'uid': ds._df['uid'].values.astype(np.int32), 'iid': ds._df['iid'].values.astype(np.int32),
           'neg_ratio': neg_ratio, 'thr': thr}
    for s in seeds:
        triples = PointSampler(ds, neg_ratio, thr, s).sample(n)
        out[f'triples_seed{s}'] = np.array([[u, i, float(v)] for u, i, v in triples], dtype=np.float64)
    ds.select_user_interaction_vec(0)           # builds the cached scipy matrices
    m = ds._cached_interaction_matrix.tocsr()
    m.sort_indices()
    out['csr_indptr'], out['csr_indices'], out['csr_data'] = m.indptr, m.indices, m.data
    mt = ds._cached_trans_interaction_matrix.tocsr()
    mt.sort_indices()
    out['csc_indptr'], out['csc_indices'], out['csc_data'] = mt.indptr, mt.indices, mt.data
    # a few explicit dense rows through the public accessor
    out['dense_user0'] = ds.select_user_interaction_vec(0).toarray().ravel()
    out['dense_item0'] = ds.select_item_interaction_vec(0).toarray().ravel()
    np.savez_compressed(os.path.join(HERE, f'sampler_{name}.npz'), **out)
    print('wrote', name, {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})


class FakeModel:
    """Deterministic stand-in exposing exactly what ranking_evaluation touches (model.rank,
    model.interaction_dataset, model.interaction_threshold, model.n_items); mirrors RecommenderABC.rank
    (recommender_abc.py:421-461) with a hash score."""

    def __init__(self, ds_train, record):
        ds_train.assign_internal_ids()
        self.interaction_dataset = ds_train
        self.interaction_threshold = 0.001
        self.n_items = ds_train.count_unique('iid')
        self.record = record

    @staticmethod
    def score(uid, iid):
        return float((uid * 7919 + iid * 104729) % 97) / 97.0

    def rank(self, user_id, item_ids, novelty=True, skip_invalid_items=True, **kwds):
        ds = self.interaction_dataset
        uid = ds.user_to_uid(user_id)
        iids = [ds.item_to_iid(it) for it in item_ids]
        iids = [i for i in iids if i is not None]
        if novelty:
            rated = set(ds.select(f'uid == {uid}').values_list('iid', to_list=True))
            cand = set(iids) - rated
        else:
            cand = set(iids)
        ranked = sorted([(self.score(uid, i), i) for i in cand], reverse=True)
        out = [(s, ds.iid_to_item(i)) for s, i in ranked]
        self.record[int(user_id)] = ([int(x) for x in item_ids], [int(it) for _, it in out])
        return out


def ranking_golden():
    rng = random.Random(0)     # the reference's fixture, test_ranking_evaluation.py:12-19
    df = pd.DataFrame([[u, i, rng.randint(-1, 5)] for u in range(50) for i in range(200) if rng.randint(0, 4) == 0],
                      columns=['user', 'item', 'interaction'])
    train, test = leave_k_out(InteractionDataset.read_df(df, verbose=False), k=5, min_user_interactions=0,
                              last_timestamps=False, seed=10, verbose=False)
    tr = train._df[['user', 'item', 'interaction']].values.tolist()
    te = test._df[['user', 'item', 'interaction']].values.tolist()
    cases = {
        'all_pos_all_neg': dict(k=2, n_pos_interactions=None, n_neg_interactions=None,
                                generate_negative_pairs=False, novelty=False),
        'gen_neg_20': dict(k=2, n_pos_interactions=None, n_neg_interactions=20, generate_negative_pairs=True,
                           novelty=False),
        'lim_neg_1': dict(k=2, n_pos_interactions=None, n_neg_interactions=1, generate_negative_pairs=False,
                          novelty=False),
        'k_list': dict(k=[1, 5, 10], n_pos_interactions=None, n_neg_interactions=None,
                       generate_negative_pairs=False, novelty=False),
        'leave1_100neg_novel': dict(k=10, n_pos_interactions=1, n_neg_interactions=100,
                                    generate_negative_pairs=True, novelty=True, seed=10),
        'pos2_float_neg': dict(k=[3, 5], n_pos_interactions=2, n_neg_interactions=1.5,
                               generate_negative_pairs=True, novelty=True, seed=3, n_test_users=30),
    }
    out = {'train_rows': tr, 'test_rows': te, 'cases': {}}
    for name, kw in cases.items():
        record = {}
        model = FakeModel(train, record)
        res = ranking_evaluation(model, test, verbose=False, max_concurrent_threads=1,
                                 metrics=[Precision(), Recall(), HitRatio(), NDCG()], **kw)
        out['cases'][name] = {'kwargs': kw, 'result': res,
                              'per_user': {str(u): {'candidates': c, 'ranked': r} for u, (c, r) in record.items()}}
        print(name, res)
    with open(os.path.join(HERE, 'ranking.json'), 'w') as f:
        json.dump(out, f)

    # --- drive the ORACLE protocol with the live UserKNN and check the reference's own KATs
    from DRecPy.Recommender.Baseline import UserKNN
    from oracle.ranking import ranking_evaluation_oracle
    knn = UserKNN(k=3, m=0, sim_metric='cosine', aggregation='weighted_mean', shrinkage=100, use_averages=False)
    knn.fit(train, verbose=False)
    train_pos = {}
    for u, it, v in tr:
        if v >= 0.001:
            train_pos.setdefault(int(u), set()).add(int(it))
    tu, ti, tv = zip(*te)

    def rank_fn(user, items, novelty):
        return [it for _, it in knn.rank(user, items, novelty=novelty, skip_invalid_items=True)]
    kat = {   # test_ranking_evaluation.py:30-60
        'test_ranking_evaluation_0': (dict(k=2), {'HitRatio@2': 0.3137, 'NDCG@2': 0.4093, 'Precision@2': 0.7021,
                                                  'Recall@2': 0.3137}),
        'test_ranking_evaluation_1': (dict(k=2, n_neg_interactions=20, generate_negative_pairs=True),
                                      {'HitRatio@2': 0.0943, 'NDCG@2': 0.1249, 'Precision@2': 0.16,
                                       'Recall@2': 0.0943}),
        'test_ranking_evaluation_2': (dict(k=2, n_neg_interactions=1),
                                      {'HitRatio@2': 0.3337, 'NDCG@2': 0.4341, 'Precision@2': 0.8111,
                                       'Recall@2': 0.3337}),
        'test_ranking_evaluation_3': (dict(k=[1, 5, 10]),
                                      {'HitRatio@1': 0.1953, 'HitRatio@10': 0.4107, 'HitRatio@5': 0.4107,
                                       'NDCG@1': 0.3968, 'NDCG@10': 0.4189, 'NDCG@5': 0.4189,
                                       'Precision@1': 0.7447, 'Precision@10': 0.7089, 'Precision@5': 0.7089,
                                       'Recall@1': 0.1953, 'Recall@10': 0.4107, 'Recall@5': 0.4107}),
    }
    check = {}
    for name, (kw, expected) in kat.items():
        got = ranking_evaluation_oracle(rank_fn, tu, ti, tv, train_pos, knn.n_items, 0.001, novelty=False, **kw)
        check[name] = {'expected': expected, 'oracle_with_live_UserKNN': got, 'equal': got == expected}
        print(name, got == expected)
    with open(os.path.join(HERE, 'ref_kat_check.json'), 'w') as f:
        json.dump(check, f, indent=1)
    assert all(c['equal'] for c in check.values()), 'oracle protocol does not reproduce the reference KATs'


def splits_golden():
    rng = np.random.default_rng(21)
    n = 4000
    user = rng.integers(0, 120, n) * 7 + 3          # raw ids out of order, not contiguous
    item = rng.integers(0, 300, n)
    key = np.unique(user * 1000 + item)
    rng.shuffle(key)
    user, item = key // 1000, key % 1000
    val = rng.integers(1, 6, len(key))
    ts = rng.integers(0, 50, len(key))               # many ties: (timestamp, rid) ordering matters
    df = pd.DataFrame({'user': user, 'item': item, 'interaction': val, 'timestamp': ts})
    cases = {
        'k3_seed10': dict(k=3, seed=10),
        'k1_seed0': dict(k=1, seed=0),
        'ratio_0.2_seed5': dict(k=0.2, seed=5),
        'k30_min25_seed7': dict(k=30, min_user_interactions=25, seed=7),     # k above some users' row counts
        'ratio_0.5_min20_seed1': dict(k=0.5, min_user_interactions=20, seed=1),
        'k2_last_timestamps': dict(k=2, last_timestamps=True, seed=3),
        'ratio_0.3_last_timestamps_min28': dict(k=0.3, last_timestamps=True, min_user_interactions=28, seed=3),
    }
    out = {'rows': df.values.tolist(), 'cases': {}}
    for name, kw in cases.items():
        ds = InteractionDataset.read_df(df, verbose=False)
        train, test = leave_k_out(ds, verbose=False, **kw)
        out['cases'][name] = {'kwargs': kw, 'train_rid': sorted(train.values_list('rid', to_list=True)),
                              'test_rid': sorted(test.values_list('rid', to_list=True))}
        print(name, len(train), len(test))
    with open(os.path.join(HERE, 'splits.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    if sys.argv[1:] == ['splits']:
        splits_golden()
        sys.exit(0)
    sampler_golden('small_zero_rows', 300, 500, 6000, data_seed=10, seeds=[10, 23, 0], n=1500, zero_frac=0.15)
    sampler_golden('small_dups', 120, 90, 2500, data_seed=11, seeds=[10], n=600, dup=200)
    sampler_golden('small_float', 200, 400, 8000, data_seed=12, seeds=[10, 7], n=800, float_vals=True)
    sampler_golden('thr3', 150, 250, 5000, data_seed=13, seeds=[10], n=600, thr=3)
    ranking_golden()
    splits_golden()
