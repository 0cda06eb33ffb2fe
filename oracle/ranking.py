"""Oracle restatement of the ranking_evaluation protocol and the metrics on the path (test infrastructure).

PINNED against the live reference (tests/golden/ranking_*.json) and the reference KATs
(tests/Evaluation/Metrics/test_ranking.py:11-99).  Follows /root/reference:
  DRecPy/Evaluation/Processes/ranking_evaluation.py:92-133   user order, per-user random.Random(seed+idx),
                                                             round(sum/count, 4)
  DRecPy/Evaluation/Processes/ranking_evaluation.py:163-246  per-user candidate generation + metric calls
  DRecPy/Evaluation/Metrics/ranking.py:32-56 (DCG), :72-91 (NDCG), :96-114 (HitRatio), Recall, Precision
The RNG is CPython's random.Random (sample / randint / shuffle), exactly what the reference runs on.
"""
import math
import random


# ---------------------------------------------------------------- metrics (ranking.py)
def dcg(recommendations, k=None, relevancies=None, strong=True):
    if relevancies is None:
        return 0
    if k is not None:
        recommendations = recommendations[:k]
    cur = 0
    for i, r in enumerate(recommendations):
        rel = float(relevancies[r])
        cur += ((2 ** rel - 1) if strong else rel) / math.log2(2 + i)
    return cur


def ndcg(recommendations, k=None, relevancies=None, strong=True):
    if relevancies is None:
        return 0
    cur = dcg(recommendations, k, relevancies, strong)
    best = sorted(relevancies.keys(), key=lambda x: -relevancies[x])
    return cur / dcg(best, k, relevancies, strong)


def hit_ratio(recommendations, k=None, relevant_recommendations=None):
    if relevant_recommendations is None:
        return 0
    if k is not None:
        recommendations = recommendations[:k]
    rec = set(str(x) for x in recommendations)
    rel = set(str(x) for x in relevant_recommendations)
    return len(rec & rel) / len(rel)


def recall(recommendations, k=None, relevant_recommendations=None):
    if relevant_recommendations is None:
        return 0
    if k is not None:
        recommendations = recommendations[:k]
    return len(set(recommendations) & set(relevant_recommendations)) / len(relevant_recommendations)


def precision(recommendations, k=None, relevant_recommendations=None):
    if relevant_recommendations is None:
        return 0
    if k is not None:
        recommendations = recommendations[:k]
    return len(set(recommendations) & set(relevant_recommendations)) / len(recommendations)


METRICS = {'Precision': ('rr', precision), 'Recall': ('rr', recall), 'HitRatio': ('rr', hit_ratio),
           'NDCG': ('rel', ndcg)}


# ---------------------------------------------------------------- protocol (ranking_evaluation.py)
def user_candidates(rng, test_rows, train_pos_items, n_items, thr, n_pos, n_neg, generate_negative_pairs,
                    train_evaluation=False):
    """ranking_evaluation.py:163-219 for one user.

    test_rows: [(raw_item, interaction)] of this user in test-set row order.
    train_pos_items: set of raw item ids with interaction >= thr for this user in the model's training set.
    Returns None if the user is skipped, else (all_items (shuffled), positives, relevancies dict)."""
    pos = [(it, v) for it, v in test_rows if v >= thr]
    if n_pos is None:
        chosen = pos
    else:
        if len(pos) < n_pos:
            return None
        chosen = rng.sample(pos, n_pos)
    positives = [it for it, _ in chosen]
    neg_pool = [it for it, v in test_rows if v < thr]
    if n_neg is None:
        negatives = list(neg_pool)
    else:
        if isinstance(n_neg, float):
            n_neg = int(n_neg * len(positives))
        negatives = rng.sample(neg_pool, min(n_neg, len(neg_pool)))
        if len(negatives) < n_neg and generate_negative_pairs:
            blacklist = set(it for it, _ in pos) if train_evaluation else \
                set(train_pos_items) | set(it for it, _ in pos)
            if n_items - len(blacklist) < n_neg:
                return None
            while len(negatives) < n_neg:
                new_item = rng.randint(0, n_items - 1)          # raw/internal id confusion kept (:211)
                if new_item not in blacklist and new_item not in negatives:
                    negatives.append(new_item)
    all_items = positives + negatives
    if len(all_items) == 0:
        return None
    rng.shuffle(all_items)
    first_val = {}
    for it, v in test_rows:                                      # select_one -> first matching row (:223)
        first_val.setdefault(it, v)
    relevancies = {it: (first_val.get(it) or 0) for it in all_items}
    return all_items, positives, relevancies


def ranking_evaluation_oracle(rank_fn, test_users, test_items, test_vals, train_pos_by_user, n_items, thr,
                              n_test_users=None, k=10, n_pos_interactions=None, n_neg_interactions=None,
                              generate_negative_pairs=False, novelty=False, seed=0,
                              metrics=('Precision', 'Recall', 'HitRatio', 'NDCG'), train_evaluation=False,
                              record=None):
    """rank_fn(user, all_items, novelty) -> ranked list of raw item ids.  Arrays are the test set in row order.
    train_pos_by_user: dict raw user -> set(raw items with interaction >= thr in the training set)."""
    ks = k if isinstance(k, list) else [k]
    sums = {(m, k_): [0, 0] for m in metrics for k_ in ks}
    order, rows = [], {}
    for u, it, v in zip(test_users, test_items, test_vals):
        if u not in rows:
            rows[u] = []
            order.append(u)
        rows[u].append((it, v))
    n_test_users = len(order) if n_test_users is None else min(n_test_users, len(order))
    for idx, user in enumerate(order[:n_test_users]):
        rng = random.Random(seed + idx)
        res = user_candidates(rng, rows[user], train_pos_by_user.get(user, set()), n_items, thr,
                              n_pos_interactions, n_neg_interactions, generate_negative_pairs, train_evaluation)
        if res is None:
            continue
        all_items, positives, relevancies = res
        recommendations = list(rank_fn(user, all_items, novelty))
        if record is not None:
            record.append((user, list(all_items), list(recommendations)))
        for m in metrics:
            kind, fn = METRICS[m]
            for k_ in ks:
                try:
                    if kind == 'rr':
                        val = fn(recommendations, k=k_, relevant_recommendations=positives)
                    else:
                        val = fn(recommendations, k=k_, relevancies=relevancies)
                    sums[(m, k_)][0] += val
                    sums[(m, k_)][1] += 1
                except Exception:
                    pass
    return {f'{m}@{k_}': (round(sums[(m, k_)][0] / sums[(m, k_)][1], 4) if sums[(m, k_)][1] > 0 else 0)
            for m, k_ in sums}
