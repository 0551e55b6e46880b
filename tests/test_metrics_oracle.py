"""The metrics oracle against the known-answer vectors in the reference's own docstrings
(replay/metrics/ndcg.py:36-48, hitrate.py, map.py, mrr.py, precision.py, recall.py doctests)."""
import pytest

from oracle import metrics_oracle as M

# pred / true of the NDCG doctest (ndcg.py:36-48): users 1 and 2
RECS = {1: [4, 5], 2: [6, 7]}
TRUE = {1: [1, 2, 3, 4, 5], 2: [8]}


def test_ndcg_doctest_vector():
    assert M.rank_metrics(RECS, TRUE, [2])["NDCG"][2] == pytest.approx(0.5, abs=1e-15)


def test_per_user_known_answers():
    # hand-checked values of the reference formulas on a small case
    pred, gt = [3, 9, 5, 7], [5, 3, 1]
    assert M.hitrate(2, pred, gt) == 1 and M.hitrate(1, [9], gt) == 0
    assert M.mrr(4, pred, gt) == 1.0 and M.mrr(4, [9, 7, 5], gt) == pytest.approx(1 / 3)
    assert M.precision(4, pred, gt) == pytest.approx(2 / 4)
    assert M.recall(4, pred, gt) == pytest.approx(2 / 3)
    assert M.map_(4, pred, gt) == pytest.approx((1 / 1 + 2 / 3) / 4)
    import math
    dcg = 1 / math.log2(2) + 1 / math.log2(4)
    idcg = 1 / math.log2(2) + 1 / math.log2(3) + 1 / math.log2(4)
    assert M.ndcg(4, pred, gt) == pytest.approx(dcg / idcg, abs=1e-15)
    # empty sides
    for fn in (M.ndcg, M.hitrate, M.map_, M.mrr, M.precision, M.recall):
        assert fn(3, [], gt) == 0
    assert M.ndcg(3, pred, []) == 0 and M.recall(3, pred, []) == 0 and M.map_(3, pred, []) == 0


def test_right_join_users_without_recs_count_as_zero():
    out = M.rank_metrics({1: [4]}, {1: [4], 2: [5]}, [1])
    assert out["HitRate"][1] == 0.5 and out["Recall"][1] == 0.5
