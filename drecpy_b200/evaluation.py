"""ranking_evaluation with batched native scoring, and the ranking metrics on the path.

Mirrors DRecPy/Evaluation/Processes/ranking_evaluation.py:19-246 (same signature, asserts, per-user
random.Random(seed + idx) candidate generation, shuffle, model.rank, relevancies, round(sum / count, 4)) and
DRecPy/Evaluation/Metrics/ranking.py (DCG :20-56, NDCG :59-91, HitRatio :94-114, Recall, Precision).
What changes: the per-user candidate generation runs in libdrb's host C++ (drb_eval_candidates) and all users
are scored in one batched GPU call (model.rank_batch) instead of 4 Python threads calling model.rank per user.
Models without rank_batch (any object with .rank) and non-integer raw ids go through the same protocol in Python.
"""
import math
import os
import random

import numpy as np
import pandas as pd

from . import _lib
from .dataset import InteractionData


# ------------------------------------------------------------------------------------------ metrics
class RankingMetricABC:
    @property
    def name(self):
        return self.__class__.__name__


class DCG(RankingMetricABC):
    def __init__(self, strong_relevancy=True):
        self.strong_relevancy = strong_relevancy

    def __call__(self, recommendations, k=None, relevancies=None):
        if relevancies is None: return 0
        if k is not None: recommendations = recommendations[:k]
        curr_dcg = 0
        for i, r in enumerate(recommendations):
            rel = float(relevancies[r])
            if self.strong_relevancy:
                curr_dcg += (2 ** rel - 1) / math.log2(2 + i)
            else:
                curr_dcg += rel / math.log2(2 + i)
        return curr_dcg


class NDCG(RankingMetricABC):
    def __init__(self, strong_relevancy=True):
        self.strong_relevancy = strong_relevancy
        self.dcg = DCG(strong_relevancy=strong_relevancy)

    def __call__(self, recommendations, k=None, relevancies=None):
        if relevancies is None: return 0
        curr_dcg = self.dcg(recommendations, relevancies=relevancies, k=k)
        best_recommendations = sorted(relevancies.keys(), key=lambda x: -relevancies[x])
        best_dcg = self.dcg(best_recommendations, relevancies=relevancies, k=k)
        return curr_dcg / best_dcg


class HitRatio(RankingMetricABC):
    def __call__(self, recommendations, k=None, relevant_recommendations=None):
        if relevant_recommendations is None: return 0
        if k is not None: recommendations = recommendations[:k]
        recommendations = set([str(item) for item in recommendations])
        relevant_recommendations = set([str(item) for item in relevant_recommendations])
        return len(recommendations.intersection(relevant_recommendations)) / len(relevant_recommendations)


class Recall(RankingMetricABC):
    def __call__(self, recommendations, k=None, relevant_recommendations=None):
        if relevant_recommendations is None: return 0
        if k is not None: recommendations = recommendations[:k]
        in_common = set(recommendations).intersection(set(relevant_recommendations))
        return len(in_common) / len(relevant_recommendations)


class Precision(RankingMetricABC):
    def __call__(self, recommendations, k=None, relevant_recommendations=None):
        if relevant_recommendations is None: return 0
        if k is not None: recommendations = recommendations[:k]
        in_common = set(recommendations).intersection(set(relevant_recommendations))
        return len(in_common) / len(recommendations)


# ------------------------------------------------------------------------------------------ candidate generation
def _group_by_user(data):
    """unique users in first-appearance order + test rows grouped by user with row order preserved."""
    codes, users = pd.factorize(data.user)
    order = np.argsort(codes, kind='stable')
    indptr = np.zeros(len(users) + 1, np.int64)
    np.cumsum(np.bincount(codes, minlength=len(users)), out=indptr[1:])
    return np.asarray(users), indptr, order


def _train_positive_items(model, users, thr):
    """What drb_eval_candidates needs to test `raw item in training positives of user` without copying anything:
    the training CSR of positives over internal ids, the training row of every evaluated user, and the sorted
    raw-item -> internal-id map (ranking_evaluation.py:196-199)."""
    data = model._data if hasattr(model, '_data') else InteractionData.from_dataset(model.interaction_dataset)
    data.assign_internal_ids()
    indptr, indices, _ = data.csr(thr)
    rows = np.ascontiguousarray(data.users_to_uids(users).astype(np.int64))
    data.items_to_iids(np.zeros(0, np.int64))          # builds the cached sorted map
    raw_sorted, order = data._cache['sorted_items']
    return (rows, np.ascontiguousarray(indptr), np.ascontiguousarray(indices),
            np.ascontiguousarray(raw_sorted.astype(np.int64)), np.ascontiguousarray(order.astype(np.int32)))


def generate_candidates(model, data, users, t_indptr, t_order, thr, n_pos, n_neg, generate_negative_pairs,
                        train_evaluation, seed):
    """ranking_evaluation.py:163-219 for every user at once (native).  Returns cand_off, cand, pos_off, pos,
    skipped."""
    n_users = len(users)
    t_item = np.ascontiguousarray(data.item[t_order].astype(np.int64))
    t_val = np.ascontiguousarray(data.interaction[t_order].astype(np.float64))
    black = None
    if not train_evaluation and generate_negative_pairs:
        black = _train_positive_items(model, users, thr)
    is_frac = isinstance(n_neg, float)
    per_user_neg = 0 if n_neg is None else (int(np.ceil(n_neg * np.diff(t_indptr).max())) if is_frac else int(n_neg))
    cap = int(len(t_item) + n_users * per_user_neg + 16)
    cand_off = np.zeros(n_users + 1, np.int64)
    pos_off = np.zeros(n_users + 1, np.int64)
    cand = np.empty(cap, np.int64)
    pos = np.empty(cap, np.int64)
    skipped = np.zeros(n_users, np.uint8)
    bp = [_lib.np_ptr(a) for a in black] if black is not None else [None] * 5
    _lib.check(_lib.load().drb_eval_candidates(
        n_users, _lib.np_ptr(t_indptr), _lib.np_ptr(t_item), _lib.np_ptr(t_val), bp[0], bp[1], bp[2], bp[3], bp[4],
        len(black[3]) if black is not None else 0, int(bool(train_evaluation)), int(model.n_items), float(thr),
        -1 if n_pos is None else int(n_pos), -1.0 if n_neg is None else float(n_neg), int(is_frac),
        int(bool(generate_negative_pairs)), int(seed), min(16, os.cpu_count() or 1), cap, _lib.np_ptr(cand_off),
        _lib.np_ptr(cand), _lib.np_ptr(pos_off), _lib.np_ptr(pos), _lib.np_ptr(skipped)))
    return cand_off, cand[:cand_off[-1]], pos_off, pos[:pos_off[-1]], skipped, t_item, t_val


def _candidates_python(rng, test_rows, train_pos_items, test_pos_items, n_items, thr, n_pos, n_neg, generate,
                       train_evaluation):
    """The same protocol in Python (non-integer raw ids)."""
    pos = [(it, v) for it, v in test_rows if v >= thr]
    if n_pos is None:
        chosen = pos
    else:
        if len(pos) < n_pos: return None
        chosen = rng.sample(pos, n_pos)
    positives = [it for it, _ in chosen]
    neg_pool = [it for it, v in test_rows if v < thr]
    if n_neg is None:
        negatives = list(neg_pool)
    else:
        if isinstance(n_neg, float): n_neg = int(n_neg * len(positives))
        negatives = rng.sample(neg_pool, min(n_neg, len(neg_pool)))
        if len(negatives) < n_neg and generate:
            blacklist = set(test_pos_items) if train_evaluation else set(train_pos_items) | set(test_pos_items)
            if n_items - len(blacklist) < n_neg: return None
            while len(negatives) < n_neg:
                new_item = rng.randint(0, n_items - 1)
                if new_item not in blacklist and new_item not in negatives:
                    negatives.append(new_item)
    all_items = positives + negatives
    if len(all_items) == 0: return None
    rng.shuffle(all_items)
    return all_items, positives


# ------------------------------------------------------------------------------------------ vectorised protocol
def _fast_evaluation(model, data, users, t_indptr, t_order, thr, n_pos, n_neg, generate, train_evaluation, seed,
                     novelty, metrics, ks):
    """Same protocol as the per-user loop below for the built-in metrics, with every per-user Python step replaced
    by array operations that perform the SAME float64 operations in the SAME order per user (sequential DCG sums,
    sequential sum over users), so the rounded results are identical.  Returns None when a corner case needs the
    general path (duplicate items inside one user's candidate / positive list)."""
    cand_off, cand, pos_off, pos, skipped, t_item, t_val = generate_candidates(
        model, data, users, t_indptr, t_order, thr, n_pos, n_neg, generate, train_evaluation, seed)
    active = np.flatnonzero(skipped == 0)
    t_item, t_val = t_item[:t_indptr[-1]], t_val[:t_indptr[-1]]    # rows of the evaluated users only
    metric_sums = {(m.name, k_): [0, 0] for m in metrics for k_ in ks}
    if len(active):
        known = model._data.users_to_uids(users[active]) >= 0       # model.rank asserts the user is known -> skipped
        active = active[known]
    if len(active):
        n = len(active)
        c_lens = (cand_off[active + 1] - cand_off[active]).astype(np.int64)
        p_lens = (pos_off[active + 1] - pos_off[active]).astype(np.int64)
        c_off = np.zeros(n + 1, np.int64); np.cumsum(c_lens, out=c_off[1:])
        p_off = np.zeros(n + 1, np.int64); np.cumsum(p_lens, out=p_off[1:])
        # equal list lengths (the usual protocol: n_pos positives + n_neg negatives for everybody) need no per-element
        # segment arithmetic: the source index is a broadcast add, the relevancy matrix a reshape
        uniform_c = int(c_lens.min()) == int(c_lens.max())
        uniform_p = int(p_lens.min()) == int(p_lens.max())
        c_seg = p_seg = None

        def flatten(off, flat, lens, tot, uniform):
            """The active users' lists back to back (+ the segment ids when the lengths differ)."""
            if uniform and len(flat) >= tot and off[active[0]] == 0 and off[active[-1] + 1] == tot:
                return flat[:tot], None        # every list in between is empty: they already lie back to back
            if uniform:
                return flat[(off[active][:, None] + np.arange(int(lens[0]), dtype=np.int64)[None, :]).ravel()], None
            seg = np.repeat(np.arange(n), lens)
            start = np.zeros(n, np.int64); np.cumsum(lens[:-1], out=start[1:])
            return flat[np.arange(tot) - np.repeat(start, lens) + np.repeat(off[active], lens)], seg
        c_flat, c_seg = flatten(cand_off, cand, c_lens, int(c_off[-1]), uniform_c)
        p_flat, p_seg = flatten(pos_off, pos, p_lens, int(p_off[-1]), uniform_p)
        big = int(max(c_flat.max(initial=0), t_item.max(initial=0), p_flat.max(initial=0))) + 2
        if big * (len(users) + 1) >= 2 ** 62 or c_flat.min(initial=0) < 0 or t_item.min(initial=0) < 0:
            return None
        # corner case: duplicate items inside one candidate or positive list -> general path.  Positives and negatives
        # are distinct test rows of the user and generated negatives avoid both (ranking_evaluation.py:163-219), so a
        # list can only repeat an item when the user's test rows do: check those (one key per test row) first.
        t_seg = np.repeat(np.arange(len(users)), np.diff(t_indptr))
        tk = np.sort(t_seg * big + t_item)
        if len(tk) > 1 and (tk[1:] == tk[:-1]).any():
            if c_seg is None: c_seg = np.repeat(np.arange(n), c_lens)
            if p_seg is None: p_seg = np.repeat(np.arange(n), p_lens)
            ck = np.sort(c_seg * big + c_flat)
            pk = np.sort(p_seg * big + p_flat)
            if (len(ck) > 1 and (ck[1:] == ck[:-1]).any()) or (len(pk) > 1 and (pk[1:] == pk[:-1]).any()):
                return None
        ranked, n_out = model.rank_arrays(users[active], c_flat, c_off, novelty=novelty)
        L = ranked.shape[1]
        # per-user DCG / ideal DCG / hit counts at every cut-off, natively for all users (drb_eval_metrics: the relevancy
        # of an item is its first test row of the user, else 0 -- ranking_evaluation.py:223 --, a hit is a ranked item
        # among the user's sampled positives); same float64 operations in the same order as the metric classes above
        lib, nthr = _lib.load(), min(16, os.cpu_count() or 1)
        t_beg = np.ascontiguousarray(t_indptr[:-1][active].astype(np.int64))
        t_end = np.ascontiguousarray(t_indptr[1:][active].astype(np.int64))
        t_item_c, t_val_c = np.ascontiguousarray(t_item, np.int64), np.ascontiguousarray(t_val, np.float64)
        ranked_c = np.ascontiguousarray(ranked, np.int64)
        n_out_c = np.ascontiguousarray(n_out, np.int32)
        c_flat_c, p_flat_c = np.ascontiguousarray(c_flat, np.int64), np.ascontiguousarray(p_flat, np.int64)
        ks_c = np.ascontiguousarray(ks, np.int64)
        dcg, idcg = np.zeros((n, len(ks))), np.zeros((n, len(ks)))
        hits_all = np.zeros((n, len(ks)), np.int64)
        _lib.check(lib.drb_eval_metrics(
            n, _lib.np_ptr(t_beg), _lib.np_ptr(t_end), _lib.np_ptr(t_item_c), _lib.np_ptr(t_val_c), _lib.np_ptr(ranked_c),
            L, _lib.np_ptr(n_out_c), _lib.np_ptr(np.ascontiguousarray(c_off[:-1])), _lib.np_ptr(np.ascontiguousarray(c_off[1:])),
            _lib.np_ptr(c_flat_c), _lib.np_ptr(np.ascontiguousarray(p_off[:-1])), _lib.np_ptr(np.ascontiguousarray(p_off[1:])),
            _lib.np_ptr(p_flat_c), _lib.np_ptr(ks_c), len(ks), nthr, _lib.np_ptr(dcg), _lib.np_ptr(idcg),
            _lib.np_ptr(hits_all)))
        for m in metrics:
            for j, k_ in enumerate(ks):
                n_rec = np.minimum(n_out, k_)
                if type(m) is NDCG:
                    cur, best = dcg[:, j], idcg[:, j]
                    good = best != 0                                # ZeroDivisionError -> metric skipped for the user
                    vals = (cur[good] / best[good]).tolist()
                else:
                    hits = hits_all[:, j]
                    if type(m) is Precision:
                        good = n_rec > 0
                        vals = (hits[good] / n_rec[good]).tolist()
                    else:                                           # HitRatio, Recall: / number of positives
                        good = p_lens > 0
                        vals = (hits[good] / p_lens[good]).tolist()
                acc = 0
                for v in vals:                                      # same left-to-right float additions as the loop
                    acc += v
                metric_sums[(m.name, k_)] = [acc, len(vals)]
    return {m + f'@{k_}': round(metric_sums[(m, k_)][0] / metric_sums[(m, k_)][1], 4)
            if metric_sums[(m, k_)][1] > 0 else 0 for m, k_ in metric_sums}


# ------------------------------------------------------------------------------------------ the protocol
def ranking_evaluation(model, ds_test=None, n_test_users=None, k=10, n_pos_interactions=None, n_neg_interactions=None,
                       generate_negative_pairs=False, novelty=False, seed=0, max_concurrent_threads=4, **kwds):
    assert n_test_users is None or n_test_users > 0, f'The number of test users ({n_test_users}) should be > 0.'
    assert n_pos_interactions is None or n_pos_interactions > 0, \
        f'The number of positive interactions ({n_pos_interactions}) should be None or an integer > 0.'
    assert n_neg_interactions is None or n_neg_interactions > 0, \
        f'The number of negative interactions ({n_neg_interactions}) should be None or an integer > 0.'
    if generate_negative_pairs and n_neg_interactions is None:
        raise Exception('Cannot generate negative interaction pairs when the number of negative interactions per user '
                        'is not defined. Either set generate_negative_pairs=False or define the n_neg_interactions '
                        'parameter.')
    interaction_threshold = kwds.get('interaction_threshold', model.interaction_threshold)
    if type(k) is not list: k = [k]
    for k_ in k: assert k_ > 0, f'k ({k_}) should be > 0.'

    train_evaluation = False
    if ds_test is None or ds_test is model.interaction_dataset or ds_test is getattr(model, '_data', None):
        train_evaluation = True
        ds_test = model.interaction_dataset
    metrics = kwds.get('metrics', [Precision(), Recall(), HitRatio(), NDCG()])
    assert isinstance(metrics, list), f'Expected "metrics" argument to be a list and found {type(metrics)}. ' \
        f'Should contain instances of RankingMetricABC.'
    for m in metrics:
        assert hasattr(m, 'name') and callable(m), f'Expected metric {m} to be an instance of type RankingMetricABC.'

    data = InteractionData.from_dataset(ds_test)
    users, t_indptr, t_order = _group_by_user(data)
    n_eval = len(users) if n_test_users is None else min(n_test_users, len(users))
    users, t_indptr = users[:n_eval], np.ascontiguousarray(t_indptr[:n_eval + 1])
    metric_sums = {(m.name, k_): [0, 0] for m in metrics for k_ in k}
    record = kwds.get('record', None)      # optional list collecting (user, candidates, ranked) for tests

    native_ids = np.issubdtype(data.item.dtype, np.integer)
    fast = (native_ids and hasattr(model, 'rank_arrays') and record is None and not kwds.get('force_python', False)
            and all(type(m) in (HitRatio, NDCG, Precision, Recall) and getattr(m, 'strong_relevancy', True)
                    for m in metrics))
    if fast:
        res = _fast_evaluation(model, data, users, t_indptr, t_order, interaction_threshold, n_pos_interactions,
                               n_neg_interactions, generate_negative_pairs, train_evaluation, seed, novelty, metrics, k)
        if res is not None:
            return res
    if native_ids:
        cand_off, cand, pos_off, pos, skipped, t_item, t_val = generate_candidates(
            model, data, users, t_indptr, t_order, interaction_threshold, n_pos_interactions, n_neg_interactions,
            generate_negative_pairs, train_evaluation, seed)
        active = np.flatnonzero(skipped == 0)
        if len(active) and hasattr(model, '_data'):
            # model.rank asserts the user is known; the reference logs that failure and skips the user
            # (ranking_evaluation.py:222 inside the per-user try) -- same filter as the vectorised path
            active = active[model._data.users_to_uids(users[active]) >= 0]
        cand_lists = [cand[cand_off[u]:cand_off[u + 1]] for u in active]
        pos_lists = [pos[pos_off[u]:pos_off[u + 1]].tolist() for u in active]
        rows = [(t_item[t_indptr[u]:t_indptr[u + 1]], t_val[t_indptr[u]:t_indptr[u + 1]]) for u in active]
    else:
        active, cand_lists, pos_lists, rows = [], [], [], []
        tr = InteractionData.from_dataset(model.interaction_dataset)
        for idx in range(n_eval):
            sl = t_order[t_indptr[idx]:t_indptr[idx + 1]]
            trows = list(zip(data.item[sl].tolist(), data.interaction[sl].tolist()))
            tpos = [it for it, v in trows if v >= interaction_threshold]
            sel = (tr.user == users[idx]) & (tr.interaction >= interaction_threshold)
            res = _candidates_python(random.Random(seed + idx), trows, tr.item[sel].tolist(), tpos, model.n_items,
                                     interaction_threshold, n_pos_interactions, n_neg_interactions,
                                     generate_negative_pairs, train_evaluation)
            if res is None: continue
            active.append(idx)
            cand_lists.append(np.asarray(res[0], dtype=object))
            pos_lists.append(res[1])
            rows.append((data.item[sl], data.interaction[sl]))

    # ---- rank: one batched call when the model supports it (model.rank semantics, ranking_evaluation.py:222)
    if len(active) == 0:
        ranked_lists = []
    elif hasattr(model, 'rank_batch') and native_ids:
        ranked_lists = []
        chunk = kwds.get('rank_chunk', 16384)
        for o in range(0, len(active), chunk):
            sub_users = users[active[o:o + chunk]].tolist()
            rl, _, _ = model.rank_batch(sub_users, cand_lists[o:o + chunk], novelty=novelty)
            ranked_lists.extend(rl)
    else:
        ranked_lists = [[item for _, item in model.rank(users[u].item() if hasattr(users[u], 'item') else users[u],
                                                         list(c.tolist()), novelty=novelty, skip_invalid_items=True)]
                        for u, c in zip(active, cand_lists)]

    # ---- metrics (ranking_evaluation.py:223-246); exceptions inside a metric skip that (metric, k) for the user
    for u, all_items, positives, (r_items, r_vals), recommendations in zip(active, cand_lists, pos_lists, rows,
                                                                          ranked_lists):
        all_items = all_items.tolist()
        first_val = {}
        for it, v in zip(r_items.tolist(), r_vals.tolist()):
            first_val.setdefault(it, v)
        relevancies = {item: (first_val.get(item) or 0) for item in all_items}
        best_item = None if len(positives) == 0 else min(positives, key=lambda it: relevancies[it])
        if record is not None:
            record.append((users[u].item() if hasattr(users[u], 'item') else users[u], all_items, list(recommendations)))
        for m in metrics:
            param_names = m.__call__.__code__.co_varnames
            for k_ in k:
                params = dict()
                for param_name in param_names:
                    if param_name == 'recommendations': params[param_name] = recommendations
                    elif param_name == 'relevant_recommendations': params[param_name] = positives
                    elif param_name == 'relevant_recommendation': params[param_name] = best_item
                    elif param_name == 'relevancies': params[param_name] = relevancies
                    elif param_name == 'k': params[param_name] = k_
                try:
                    metric_sums[(m.name, k_)][0] += m(**params)
                    metric_sums[(m.name, k_)][1] += 1
                except Exception:
                    pass

    return {m + f'@{k_}': round(metric_sums[(m, k_)][0] / metric_sums[(m, k_)][1], 4)
            if metric_sums[(m, k_)][1] > 0 else 0 for m, k_ in metric_sums}
