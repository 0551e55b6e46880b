"""GPU ranking metrics (cql_rank_metrics) against the metrics oracle."""
import numpy as np
import pytest

from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("U,k_rec,n_items", [(1, 2, 10), (257, 10, 300), (1000, 37, 5000)])
def test_rank_metrics_match_oracle(engine_factory, U, k_rec, n_items):
    eng = engine_factory(batch_size=64)
    rng = np.random.default_rng(5)
    n_ids = 3 * U + 5
    users = rng.choice(n_ids, size=U, replace=False).astype(np.int32)
    rec = np.full((U, k_rec), -1, dtype=np.int32)
    recs, gt = {}, {}
    for r, u in enumerate(users):
        n_pred = int(rng.integers(0, k_rec + 1)) if r % 5 else k_rec        # ragged, some empty
        row = rng.choice(n_items, size=n_pred, replace=False)
        rec[r, :n_pred] = row
        recs[int(u)] = row.tolist()
        n_gt = int(rng.integers(0, 40)) if r % 7 else 0                      # some users without ground truth
        truth = rng.choice(n_items, size=n_gt, replace=False)
        m = min(n_gt, n_pred, 3)
        if m and r % 2:                                                      # make hits likely
            truth[:m] = row[:m]
        gt[int(u)] = sorted(set(truth.tolist()))
    indptr = np.zeros(n_ids + 1, dtype=np.int64)
    flat = []
    for uid in range(n_ids):
        flat.extend(gt.get(uid, []))
        indptr[uid + 1] = len(flat)
    ks = [1, 5, 10, 50]
    got = eng.rank_metrics(rec, users, indptr, np.asarray(flat, dtype=np.int32), ks)
    ref = M.rank_metrics(recs, {int(u): gt[int(u)] for u in users}, ks)
    for name in ref:
        for k in ks:
            assert got[name][k] == pytest.approx(ref[name][k], rel=1e-12, abs=1e-15), (name, k)


def test_rank_metrics_argument_errors(engine_factory):
    eng = engine_factory(batch_size=64)
    with pytest.raises(ValueError):
        eng.rank_metrics(np.zeros((2, 3), np.int32), [1], np.zeros(4, np.int64), [], [1])
    with pytest.raises(ValueError):
        eng.rank_metrics(np.zeros((1, 3), np.int32), [1], np.zeros(3, np.int64), [], [0])
    out = eng.rank_metrics(np.zeros((0, 3), np.int32), [], np.zeros(1, np.int64), [], [1, 2])
    assert out["NDCG"][1] == 0.0
