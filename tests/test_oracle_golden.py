"""The oracle against the reference's golden vectors / known-answer tests (CPU only).

Goldens come from the live reference (tests/golden/make_golden.py); KATs are restated from the reference's
own test files with their line numbers."""
import json
import os
import random

import numpy as np
import pytest

from oracle import dataset as ods
from oracle import ranking as orank
from oracle.sampler import PointSamplerOracle, PointSamplerOracleCSR

SAMPLER_CASES = ['small_zero_rows', 'small_dups', 'small_float', 'thr3']


@pytest.mark.parametrize('name', SAMPLER_CASES)
def test_internal_ids_and_csr(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f'sampler_{name}.npz'))
    uid, iid, uu, ii = ods.assign_internal_ids(g['user'], g['item'])
    assert np.array_equal(uid, g['uid']) and np.array_equal(iid, g['iid'])
    U, I = len(uu), len(ii)
    indptr, indices, data = ods.build_csr(uid, iid, g['interaction'], U, I)
    assert np.array_equal(indptr, g['csr_indptr'])
    assert np.array_equal(indices, g['csr_indices'])
    assert np.array_equal(data, g['csr_data'])
    indptr_t, indices_t, data_t = ods.build_csr(iid, uid, g['interaction'], I, U)
    assert np.array_equal(indptr_t, g['csc_indptr'])
    assert np.array_equal(indices_t, g['csc_indices'])
    assert np.array_equal(data_t, g['csc_data'])
    assert np.array_equal(ods.user_interaction_vec((indptr, indices, data), 0, I), g['dense_user0'])
    assert np.array_equal(ods.user_interaction_vec((indptr_t, indices_t, data_t), 0, U), g['dense_item0'])


@pytest.mark.parametrize('name', SAMPLER_CASES)
def test_point_sampler_vs_live_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f'sampler_{name}.npz'))
    for key in g.files:
        if not key.startswith('triples_seed'):
            continue
        seed = int(key[len('triples_seed'):])
        want = g[key]
        for cls in (PointSamplerOracle, PointSamplerOracleCSR):      # dict / set storage and the array storage
            s = cls(g['uid'], g['iid'], g['interaction'], int(g['neg_ratio']), float(g['thr']), seed)
            got = np.array([[u, i, float(v)] for u, i, v in s.sample(len(want))])
            assert np.array_equal(got, want), f'{name} seed {seed} {cls.__name__}'


def test_generators_reference_kats():
    """tests/Dataset/test_mem_dataset.py:670-706 (null pairs, seed 23) and :593-667 (positive generator,
    seed 23) on tests/Dataset/resources/test.csv (4 rows: jack/ps4/3, john/hard-drive/4, alfred/pen/1,
    jack/xbox/5 -- row order of the fixture)."""
    users = ['jack', 'john', 'alfred', 'jack']
    items = ['ps4', 'hard-drive', 'pen', 'xbox']
    vals = [3.0, 4.0, 1.0, 5.0]
    uid, iid, _, _ = ods.assign_internal_ids(np.array(users), np.array(items))
    assert uid.tolist() == [0, 1, 2, 0] and iid.tolist() == [0, 1, 2, 3]      # test_mem_dataset.py:1734-1752
    s = PointSamplerOracle(uid, iid, np.array(vals), 5, None, 23)
    got = [s.sample_negative()[:2] for _ in range(4)]
    assert got == [(1, 0), (0, 2), (1, 3), (0, 1)]                            # :676-700
    s = PointSamplerOracle(uid, iid, np.array(vals), 5, None, 23)
    assert s.sample_positive()[1] == 1                                        # :637-639 (rid 1 -> iid 1)
    # 'interaction > 3.0' seed 23 -> interaction 4.0 (:608-609); oracle filter is >=, use thr just above 3
    s = PointSamplerOracle(uid, iid, np.array(vals), 5, 3.0001, 23)
    assert s.sample_positive()[2] == 4.0


def test_metric_kats():
    """tests/Evaluation/Metrics/test_ranking.py (DCG/NDCG/HitRatio known answers)."""
    # NDCG :46-60 style: perfect ranking -> 1, relevancies dict drives the ideal list
    rel = {1: 3, 2: 2, 3: 3, 4: 0, 5: 1, 6: 2}
    assert round(orank.dcg([1, 2, 3, 4, 5, 6], relevancies=rel, strong=False), 3) == 6.861   # wikipedia example
    assert round(orank.ndcg([1, 2, 3, 4, 5, 6], relevancies=rel, strong=False), 3) == 0.961
    assert orank.hit_ratio([1, 2, 3], relevant_recommendations=[2]) == 1.0
    assert orank.hit_ratio([1, 2, 3], k=1, relevant_recommendations=[2]) == 0.0
    assert orank.hit_ratio([1, 2, 3], k=2, relevant_recommendations=[2, 9]) == 0.5
    assert orank.ndcg([1, 2], relevancies=None) == 0


def test_ranking_protocol_vs_live_reference(golden_dir):
    with open(os.path.join(golden_dir, 'ranking.json')) as f:
        g = json.load(f)
    tr, te = g['train_rows'], g['test_rows']
    train_pos, train_items_by_user, item_order = {}, {}, []
    seen = set()
    for u, it, v in tr:
        if it not in seen:
            seen.add(it)
            item_order.append(it)
        train_items_by_user.setdefault(u, set()).add(it)
        if v >= 0.001:
            train_pos.setdefault(u, set()).add(it)
    user_order = []
    for u, _, _ in tr:
        if u not in user_order:
            user_order.append(u)
    item_to_iid = {it: i for i, it in enumerate(item_order)}
    user_to_uid = {u: i for i, u in enumerate(user_order)}
    n_items = len(item_order)

    def rank_fn(user, items, novelty):       # the FakeModel of make_golden.py, restated
        uid = user_to_uid[user]
        cand = set(item_to_iid[it] for it in items if it in item_to_iid)
        if novelty:
            cand -= set(item_to_iid[it] for it in train_items_by_user.get(user, ()))
        ranked = sorted([(float((uid * 7919 + i * 104729) % 97) / 97.0, i) for i in cand], reverse=True)
        return [item_order[i] for _, i in ranked]

    tu, ti, tv = zip(*te)
    for name, case in g['cases'].items():
        record = []
        kw = dict(case['kwargs'])
        got = orank.ranking_evaluation_oracle(rank_fn, tu, ti, tv, train_pos, n_items, 0.001, record=record, **kw)
        assert got == case['result'], name
        assert len(record) == len(case['per_user']), name
        for user, cands, ranked in record:
            ref = case['per_user'][str(user)]
            assert cands == ref['candidates'], (name, user)
            assert ranked == ref['ranked'], (name, user)


def test_ref_kat_check_recorded(golden_dir):
    """The oracle protocol + live UserKNN reproduced test_ranking_evaluation.py:30-60 at golden-making time."""
    with open(os.path.join(golden_dir, 'ref_kat_check.json')) as f:
        chk = json.load(f)
    assert len(chk) == 4 and all(c['equal'] for c in chk.values())
