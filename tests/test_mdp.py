"""CPU: the vectorised MDP builder against the loop oracle; edge cases."""
import numpy as np
import pandas as pd
import pytest

from oracle import mdp_oracle
from replay_cql_b200.mdp import build_mdp, seen_csr, to_transitions
from replay_cql_b200.synthetic import make_log


def _check(log, top_k, seed=0):
    noise = np.random.default_rng(seed).standard_normal(len(log)) * 1e-3
    got = to_transitions(build_mdp(log, top_k=top_k, action_noise=noise))
    ref = mdp_oracle.build_mdp(log, top_k=top_k, action_noise=noise)
    for k in ref:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)


@pytest.mark.parametrize("top_k", [0, 1, 3, 1000])
def test_random_log_with_ties(top_k):
    rng = np.random.default_rng(1)
    n = 500
    log = pd.DataFrame({"user_idx": rng.integers(0, 23, n), "item_idx": rng.integers(0, 50, n),
                        "timestamp": rng.integers(0, 40, n),        # many timestamp ties
                        "relevance": rng.integers(1, 6, n).astype(float)})
    _check(log, top_k)


def test_datetime_timestamps_and_shuffled_rows():
    log = make_log("tiny").sample(frac=1.0, random_state=3).reset_index(drop=True)
    log["timestamp"] = pd.to_datetime(log["timestamp"], unit="s")
    _check(log, 10)


def test_single_row_and_single_user():
    _check(pd.DataFrame({"user_idx": [7], "item_idx": [3], "timestamp": [1], "relevance": [5.0]}), 1)
    _check(pd.DataFrame({"user_idx": [2] * 5, "item_idx": [1, 2, 3, 4, 5], "timestamp": [5, 4, 3, 2, 1],
                         "relevance": [1.0, 5.0, 5.0, 2.0, 3.0]}), 2)


def test_empty_log():
    m = build_mdp(pd.DataFrame({"user_idx": [], "item_idx": [], "timestamp": [], "relevance": []}))
    assert len(m) == 0 and m.obs.shape == (0, 2)


def test_rejects_indices_not_exact_in_float32():
    with pytest.raises(ValueError):
        build_mdp(pd.DataFrame({"user_idx": [2 ** 24], "item_idx": [0], "timestamp": [0], "relevance": [1.0]}))
    with pytest.raises(ValueError):
        build_mdp(pd.DataFrame({"user_idx": [1], "item_idx": [0]}))


def test_seen_csr_sorted_unique():
    log = pd.DataFrame({"user_idx": [3, 1, 1, 3, 3], "item_idx": [9, 4, 4, 2, 9]})
    indptr, items = seen_csr(log, 5)
    assert indptr.tolist() == [0, 0, 1, 1, 3, 3]
    assert items.tolist() == [4, 2, 9]


def test_synthetic_shapes():
    log = make_log("tiny")
    assert len(log) == 4000 and log["user_idx"].nunique() == 64
    assert log.groupby("user_idx")["timestamp"].apply(lambda s: (np.diff(s.to_numpy()) > 0).all()).all()
