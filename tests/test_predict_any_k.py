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
