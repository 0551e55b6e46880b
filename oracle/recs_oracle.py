"""CPU restatement of the reference's predict post-processing.  TEST INFRASTRUCTURE ONLY.

These ARE in the mounted checkout and are restated line by line (on pandas,
because no JVM/pyspark exists here):

* ``filter_seen``   <- ``replay/models/base_rec.py:417-464`` (``_filter_seen``)
* ``top_k_recs``    <- ``replay/utils.py:100-127`` (``get_top_k`` / ``get_top_k_recs``)
* ``predict_wrap``  <- ``replay/models/base_rec.py:467-539`` (``_predict_wrap``) for a
  model whose ``_predict`` returns ALL user x item pairs (Appendix B behaviour).

Spark's ``row_number`` over a single sort key breaks ties arbitrarily; this
restatement breaks them by ascending item_idx so results are reproducible.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _rank_desc(recs: pd.DataFrame) -> pd.Series:
    order = recs.sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True], kind="stable")
    rank = order.groupby("user_idx").cumcount() + 1
    return rank.reindex(recs.index)


def top_k_recs(recs: pd.DataFrame, k: int) -> pd.DataFrame:
    """base_rec.py:526 -> utils.py:112-127: row_number() over (user order by relevance desc) <= k."""
    if len(recs) == 0:
        return recs[["user_idx", "item_idx", "relevance"]]
    r = _rank_desc(recs)
    return recs[r <= k][["user_idx", "item_idx", "relevance"]]


def filter_seen(recs: pd.DataFrame, log: pd.DataFrame, k: int, users: pd.DataFrame) -> pd.DataFrame:
    """base_rec.py:417-464."""
    users_log = log.merge(users[["user_idx"]].drop_duplicates(), on="user_idx")       # :424
    num_seen = users_log.groupby("user_idx")["item_idx"].count().rename("seen_count").reset_index()  # :426-428
    max_seen = int(num_seen["seen_count"].max()) if len(num_seen) else 0                 # :431-434
    recs = recs.copy()
    recs["temp_rank"] = _rank_desc(recs) if len(recs) else []                            # :437-444
    recs = recs[recs["temp_rank"] <= max_seen + k]
    recs = recs.merge(num_seen, on="user_idx", how="left").fillna({"seen_count": 0})     # :447-452
    recs = recs[recs["temp_rank"] <= recs["seen_count"] + k].drop(columns=["temp_rank", "seen_count"])
    seen_pairs = set(zip(users_log["user_idx"].tolist(), users_log["item_idx"].tolist()))  # :455-462 anti-join
    keep = [(u, i) not in seen_pairs for u, i in zip(recs["user_idx"].tolist(), recs["item_idx"].tolist())]
    return recs[np.asarray(keep, dtype=bool)] if len(recs) else recs


def predict_wrap(all_pairs: pd.DataFrame, log: pd.DataFrame | None, k: int,
                 users: pd.DataFrame, filter_seen_items: bool = True) -> pd.DataFrame:
    """base_rec.py:514-528 given the model's full user x item relevance frame."""
    recs = all_pairs
    if filter_seen_items and log is not None:
        recs = filter_seen(recs, log, k, users)
    out = top_k_recs(recs, k)
    return out.sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True]).reset_index(drop=True)


def brute_force_topk(score_fn, users: np.ndarray, items: np.ndarray, seen: dict, k: int):
    """Per-user loop in the style of ``replay/models/neuromf.py:394-438``: build the
    I x 2 observation block, one forward per user, drop seen, keep k.
    ``score_fn(obs[I,2]) -> relevance[I]``; ``seen``: user -> set(items).
    Returns (items[U,k] int32 padded with -1, scores[U,k] float32 padded with -inf).
    """
    U = len(users)
    out_i = np.full((U, k), -1, dtype=np.int32)
    out_s = np.full((U, k), -np.inf, dtype=np.float32)
    for r, u in enumerate(users):
        obs = np.stack([np.full(len(items), float(u), dtype=np.float32), items.astype(np.float32)], axis=1)
        sc = np.asarray(score_fn(obs), dtype=np.float32)
        mask = np.array([int(i) not in seen.get(int(u), ()) for i in items], dtype=bool)
        cand_i, cand_s = items[mask], sc[mask]
        order = np.lexsort((cand_i, -cand_s.astype(np.float64)))[:k]
        out_i[r, :len(order)] = cand_i[order]
        out_s[r, :len(order)] = cand_s[order]
    return out_i, out_s
