"""CPU test of ``CQL._score_topk_any_k`` (models.py): the reference's ``predict`` accepts any ``k``
(``replay/models/base_rec.py:466-539``); the kernel selects at most ``MAX_TOPK`` per pass, larger ``k`` runs in passes that
treat the items already picked as seen.  The engine is replaced by a brute-force stand-in, ``MAX_TOPK`` is lowered so
that several passes are needed."""
import numpy as np
import pytest

from replay_cql_b200 import _lib
from replay_cql_b200.models import CQL


class _BruteEngine:
    """score_topk with the engine's contract: top-k unseen items per user, (-1, -inf) padded, best first."""

    def __init__(self, scores):
        self.scores = scores
        self.calls = 0

    def score_topk(self, users, items, k, seen_indptr=None, seen_items=None, mode="q"):
        assert 1 <= k <= _lib.MAX_TOPK
        self.calls += 1
        out_i = np.full((len(users), k), -1, dtype=np.int32)
        out_s = np.full((len(users), k), -np.inf, dtype=np.float32)
        for r, u in enumerate(users):
            sc = self.scores[u, items].astype(np.float32).copy()
            if seen_indptr is not None:
                seen = seen_items[seen_indptr[u]:seen_indptr[u + 1]]
                sc[np.isin(items, seen)] = -np.inf
            order = np.argsort(-sc, kind="stable")[:k]
            order = order[np.isfinite(sc[order])]
            out_i[r, :len(order)] = items[order]
            out_s[r, :len(order)] = sc[order]
        return out_i, out_s


@pytest.mark.parametrize("k,with_seen", [(3, True), (4, False), (5, True), (11, True), (40, False)])
def test_any_k_equals_direct_topk(monkeypatch, k, with_seen):
    monkeypatch.setattr(_lib, "MAX_TOPK", 4)
    rng = np.random.default_rng(k)
    n_users, n_items = 6, 40
    scores = rng.normal(size=(n_users, n_items)).astype(np.float32)
    users = np.array([0, 2, 5], dtype=np.int32)
    items = np.arange(n_items, dtype=np.int32)
    indptr = seen = None
    if with_seen:
        lists = [np.sort(rng.choice(n_items, size=rng.integers(0, 12), replace=False)).astype(np.int32) for _ in range(n_users)]
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum([len(x) for x in lists], out=indptr[1:])
        seen = np.concatenate(lists).astype(np.int32)
    model = CQL.__new__(CQL)
    model.score = "q"
    model.engine = _BruteEngine(scores)
    kk = min(k, n_items)
    ti, ts = model._score_topk_any_k(users, items, kk, indptr, seen, n_users)
    assert ti.shape == (3, kk) and model.engine.calls == -(-kk // 4)
    for r, u in enumerate(users):
        sc = scores[u].copy()
        if with_seen:
            sc[seen[indptr[u]:indptr[u + 1]]] = -np.inf
        order = np.argsort(-sc, kind="stable")
        order = order[np.isfinite(sc[order])][:kk]
        got = ti[r][ti[r] >= 0]
        assert got.tolist() == order.tolist()                 # same items, same (descending) order
        np.testing.assert_array_equal(ts[r][: len(order)], sc[order])
        assert np.all(ti[r][len(order):] == -1)               # fewer unseen items than k: padded, never repeated


def test_is_top_k_detects_ranked_output():
    """`Recommender._is_top_k`: the O(n) check that lets `_predict_wrap` skip `get_top_k_recs` for a model whose
    `_predict` already returns the k best per user, best first (ties by ascending item): it must agree with actually
    running `get_top_k_recs` -- True iff that call would return the frame unchanged."""
    import pandas as pd
    from replay_cql_b200.recommender import Recommender, get_top_k_recs, REC_COLUMNS
    rng = np.random.default_rng(0)

    def frame(u, i, r):
        return pd.DataFrame({"user_idx": np.asarray(u, np.int32), "item_idx": np.asarray(i, np.int32),
                             "relevance": np.asarray(r, np.float64)})

    good = frame([0, 0, 0, 2, 2, 5], [7, 3, 9, 1, 4, 2], [0.9, 0.5, 0.5, 2.0, 1.0, 0.1])
    cases = {
        "ranked": (good, 3, True),
        "too many rows for k": (good, 2, False),
        "relevance ascending inside a user": (frame([0, 0], [1, 2], [0.1, 0.2]), 5, False),
        "tie with descending item": (frame([0, 0], [5, 2], [0.3, 0.3]), 5, False),
        "users not grouped": (frame([0, 1, 0], [1, 2, 3], [0.9, 0.8, 0.7]), 5, False),
        "users descending": (frame([3, 1], [1, 2], [0.9, 0.8]), 5, False),
        "empty": (frame([], [], []), 3, True),
    }
    for name, (df, k, want) in cases.items():
        assert Recommender._is_top_k(df, k) is want, name
        same = get_top_k_recs(df, k)[REC_COLUMNS].reset_index(drop=True).equals(df[REC_COLUMNS].reset_index(drop=True)) if len(df) else True
        assert same is want, name
    for _ in range(50):                                        # random frames: the check never claims more than the sort shows
        n = int(rng.integers(1, 12))
        df = frame(np.sort(rng.integers(0, 4, n)), rng.integers(0, 6, n), rng.integers(0, 3, n) / 2.0)
        k = int(rng.integers(1, 5))
        if Recommender._is_top_k(df, k):
            assert get_top_k_recs(df, k)[REC_COLUMNS].reset_index(drop=True).equals(df[REC_COLUMNS].reset_index(drop=True))
