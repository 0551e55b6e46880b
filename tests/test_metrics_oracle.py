"""The metrics oracle against the known-answer vectors in the reference's own docstrings
(replay/metrics/ndcg.py:36-48, hitrate.py, map.py, mrr.py, precision.py, recall.py doctests)."""
import numpy as np
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


# ---- the reference's own golden values: /root/reference/tests/test_metrics.py (fixtures :45-95, expectations :178-305) ----
# recs fixture (:45-60) sorted by relevance per user; true fixture (:79-92) as item sets; true_users fixture (:95-100)
REF_RECS = {0: [0, 1, 2], 1: [1, 0, 4], 2: [0, 3, 2]}
REF_TRUE = {0: [0, 4, 1], 1: [5, 0], 2: [1]}
# ground_truth_users = [1, 2, 3, 4]: preprocess_gt right-joins on them (base_metric.py:73-97) -- user 0 drops out,
# users 3 and 4 enter with an empty ground truth and count as zero
REF_TRUE_USERS = {1: [5, 0], 2: [1], 3: [], 4: []}
_L = np.log2


@pytest.mark.parametrize("gt_users", [False, True])
def test_reference_golden_values(gt_users):
    out = M.rank_metrics(REF_RECS, REF_TRUE_USERS if gt_users else REF_TRUE, [1, 3])
    if not gt_users:
        want = {
            "HitRate": {3: 2 / 3, 1: 1 / 3},                                                      # test_metrics.py:180
            "NDCG": {1: 1 / 3,                                                                    # :229-241
                     3: 1 / 3 * (1 / (1 / _L(2) + 1 / _L(3) + 1 / _L(4)) * (1 / _L(2) + 1 / _L(3))
                                 + 1 / (1 / _L(2) + 1 / _L(3)) * (1 / _L(3)))},
            "Precision": {1: 1 / 3, 3: (2 / 3 + 1 / 3) / 3},                                      # :262
            "MAP": {1: 1 / 3, 3: ((1 + 1) / 3 + (0 + 1 / 2) / 3) / 3},                            # :279
            "Recall": {1: 1 / 9, 3: (1 / 2 + 2 / 3) / 3},                                         # :295
        }
    else:
        want = {
            "HitRate": {3: 1 / 4, 1: 0.0},                                                        # :180
            "NDCG": {1: 0.0, 3: 1 / 4 * (1 / (1 / _L(2) + 1 / _L(3)) * (1 / _L(3)))},             # :243-250
            "Precision": {3: 1 / 4 * 1 / 3, 1: 0.0},                                              # :263
            "MAP": {1: 0.0, 3: 1 / 2 * 1 / 3 * 1 / 4},                                            # :281
            "Recall": {1: 0.0, 3: 1 / 2 * 1 / 4},                                                 # :296
        }
    for name, per_k in want.items():
        for k, v in per_k.items():
            assert out[name][k] == pytest.approx(v, abs=1e-12), (name, k)


def test_reference_mrr_doctest_and_size_mismatch_cases():
    # mrr.py:12-17 doctest: pred [3, 2, 1], true {2, 4, 5}
    assert M.rank_metrics({1: [3, 2, 1]}, {1: [2, 4, 5]}, [3, 1])["MRR"] == {3: 0.5, 1: 0.0}
    # test_metrics.py:168-175: every quality metric is 0.5 when the test set has one more user, 1.0 the other way round
    one_user, two_users = {1: [1]}, {1: [1], 2: [2]}
    for name, per_k in M.rank_metrics(one_user, two_users, [1]).items():
        assert per_k[1] == 0.5, name
    for name, per_k in M.rank_metrics(two_users, one_user, [1]).items():
        assert per_k[1] == 1.0, name


def test_reference_per_user_edge_cases():
    # test_metrics.py:373-406 (test_empty_recs, test_bad_recs, test_not_full_recs)
    for name, fn in M.METRICS.items():
        assert fn(4, [], [2, 4]) == 0, name
        assert fn(4, [1, 3], [2, 4]) == 0, name
        if name not in ("Precision", "MAP"):
            assert fn(4, [4, 1, 2], [2, 4]) == pytest.approx(fn(3, [4, 1, 2], [2, 4]), abs=1e-15), name
