"""CPU restatement of the reference's ranking metrics.  TEST INFRASTRUCTURE ONLY.

Per-user bodies restated from the mounted checkout (``_get_metric_value_by_user``):

* NDCG       <- ``replay/metrics/ndcg.py:51-61``
* HitRate    <- ``replay/metrics/hitrate.py`` (last method)
* MAP        <- ``replay/metrics/map.py`` (last method)
* MRR        <- ``replay/metrics/mrr.py`` (last method)
* Precision  <- ``replay/metrics/precision.py`` (last method)
* Recall     <- ``replay/metrics/recall.py`` (last method)

and the user set of ``get_enriched_recommendations`` (``replay/metrics/base_metric.py:102-140``): a RIGHT join on
the ground-truth users, missing predictions filled with an empty list; the metric is the mean over those users
(``base_metric.py`` ``_mean``).  With ``ground_truth_users`` the reference right-joins the ground truth on those users
(``preprocess_gt``, ``base_metric.py:73-97``): callers model that by passing exactly those users as keys, with an empty
collection for a user without test items.  Pinned by the doctest vectors of the reference files and by the golden values
of the reference's ``tests/test_metrics.py`` (tests/test_metrics_oracle.py).
"""
from __future__ import annotations

import math

import numpy as np


def ndcg(k, pred, ground_truth) -> float:
    if len(pred) == 0 or len(ground_truth) == 0:
        return 0.0
    pred_len = min(k, len(pred))
    ground_truth_len = min(k, len(ground_truth))
    denom = [1 / math.log2(i + 2) for i in range(k)]
    dcg = sum(denom[i] for i in range(pred_len) if pred[i] in ground_truth)
    idcg = sum(denom[:ground_truth_len])
    return dcg / idcg


def hitrate(k, pred, ground_truth) -> float:
    for i in pred[:k]:
        if i in ground_truth:
            return 1
    return 0


def map_(k, pred, ground_truth) -> float:
    length = min(k, len(pred))
    if len(ground_truth) == 0 or len(pred) == 0:
        return 0
    tp_cum = 0
    result = 0
    for i in range(length):
        if pred[i] in ground_truth:
            tp_cum += 1
            result += tp_cum / (i + 1)
    return result / k


def mrr(k, pred, ground_truth) -> float:
    for i in range(min(k, len(pred))):
        if pred[i] in ground_truth:
            return 1 / (1 + i)
    return 0


def precision(k, pred, ground_truth) -> float:
    if len(pred) == 0:
        return 0
    return len(set(pred[:k]) & set(ground_truth)) / k


def recall(k, pred, ground_truth) -> float:
    if len(ground_truth) == 0:
        return 0.0
    return len(set(pred[:k]) & set(ground_truth)) / len(ground_truth)


METRICS = {"NDCG": ndcg, "HitRate": hitrate, "MAP": map_, "MRR": mrr, "Precision": precision, "Recall": recall}


def rank_metrics(recs: dict, ground_truth: dict, ks) -> dict:
    """recs: user -> list of item ids, best first; ground_truth: user -> collection of item ids.
    Mean over the ground-truth users (right join), per metric and cut-off."""
    users = list(ground_truth.keys())
    out = {}
    for name, fn in METRICS.items():
        out[name] = {}
        for k in ks:
            vals = [float(fn(k, list(recs.get(u, [])), list(ground_truth[u]))) for u in users]
            out[name][k] = float(np.mean(vals)) if vals else 0.0
    return out
