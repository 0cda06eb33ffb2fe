"""Oracle restatement of the dataset members on the hot path (test infrastructure, see oracle/__init__.py).

Follows /root/reference/DRecPy/Dataset/mem_dataset.py:
  assign_internal_ids        :309-330   (ids = order of first appearance, pd.Categorical codes)
  _build_interaction_matrix  :480-498   (scipy CSR, duplicates summed, float64 values; transpose for items)
  select_user_interaction_vec / select_item_interaction_vec  :165-218
"""
import numpy as np


def assign_internal_ids(users, items):
    """mem_dataset.py:309-330 -- internal id = rank of first appearance in row order.

    Returns (uid[n], iid[n], unique_users, unique_items) with unique_* indexed by internal id."""
    users = np.asarray(users)
    items = np.asarray(items)
    uu, first_u = np.unique(users, return_index=True)
    order_u = np.argsort(first_u, kind='stable')
    unique_users = uu[order_u]
    rank_u = np.empty(len(uu), dtype=np.int64)
    rank_u[order_u] = np.arange(len(uu))
    uid = rank_u[np.searchsorted(uu, users)]

    ii, first_i = np.unique(items, return_index=True)
    order_i = np.argsort(first_i, kind='stable')
    unique_items = ii[order_i]
    rank_i = np.empty(len(ii), dtype=np.int64)
    rank_i[order_i] = np.arange(len(ii))
    iid = rank_i[np.searchsorted(ii, items)]
    return uid.astype(np.int32), iid.astype(np.int32), unique_users, unique_items


def build_csr(rows, cols, vals, n_rows, n_cols):
    """mem_dataset.py:480-498 -- csr_matrix((interactions, (users, cols))): duplicates are summed
    (float64), column indices sorted.  Returns (indptr int64[n_rows+1], indices int32[nnz], data float64[nnz])."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    key = rows * n_cols + cols
    order = np.argsort(key, kind='stable')
    key_s = key[order]
    val_s = vals[order]
    if len(key_s) == 0:
        return np.zeros(n_rows + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float64)
    starts = np.flatnonzero(np.concatenate(([True], key_s[1:] != key_s[:-1])))
    ukey = key_s[starts]
    data = np.add.reduceat(val_s, starts)
    r = ukey // n_cols
    c = (ukey % n_cols).astype(np.int32)
    indptr = np.zeros(n_rows + 1, np.int64)
    np.add.at(indptr, r + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, c, data


def user_interaction_vec(csr, uid, n_cols):
    """select_user_interaction_vec(uid).toarray().ravel()  (mem_dataset.py:165-176)"""
    indptr, indices, data = csr
    out = np.zeros(n_cols, np.float64)
    out[indices[indptr[uid]:indptr[uid + 1]]] = data[indptr[uid]:indptr[uid + 1]]
    return out
