"""CPU: ``Recommender.optimize`` of the pandas mirror, re-targeted from the reference's tests/test_optuna.py
(:12-67) with a stand-in for ALSWrap (same ``_search_space`` shape: ``rank`` is ``loguniform_int``), plus the
search-space plumbing of ``replay_cql_b200.models.CQL`` that needs no GPU."""
import pickle

import numpy as np
import pandas as pd
import pytest

from oracle import metrics_oracle
from replay_cql_b200.recommender import RandomSearchStudy, ndcg_at_k
from tests.test_recommender_conformance import LOG, PopLike


class RankRec(PopLike):
    """PopLike whose relevance is sharpened by ``rank`` and shifted by ``beta`` -- enough for trials to differ."""

    _search_space = {"rank": {"type": "loguniform_int", "args": [8, 256]},          # ALSWrap's space (replay/models/als.py)
                     "beta": {"type": "uniform", "args": [0.0, 1.0]},
                     "kind": {"type": "categorical", "args": ["a", "b"]}}

    def __init__(self, rank: int = 10, beta: float = 0.5, kind: str = "a"):
        self.rank, self.beta, self.kind = rank, beta, kind
        self.fits = 0

    @property
    def _init_args(self):
        return {"rank": self.rank, "beta": self.beta, "kind": self.kind}

    def _fit(self, log, user_features=None, item_features=None):
        super()._fit(log)
        self.fits += 1

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        out = super()._predict(log, k, users, items)
        out["relevance"] = out["relevance"] * self.rank + self.beta * (out["item_idx"] % 2) + (self.kind == "b")
        return out


@pytest.fixture
def model():
    return RankRec()


@pytest.mark.parametrize("borders", [{"wrong_name": None}, {"rank": None}, {"rank": 2}, {"rank": [1]}, {"rank": [1, 2, 3]}],
                         ids=["wrong name", "None border", "int border", "border's too short", "border's too long"])
def test_bad_borders(model, borders):   # test_optuna.py:13-32
    with pytest.raises(ValueError):
        model._prepare_param_borders(borders)


@pytest.mark.parametrize("borders", [None, {"rank": [5, 9]}])
def test_correct_borders(model, borders):   # test_optuna.py:35-41
    res = model._prepare_param_borders(borders)
    assert res.keys() == model._search_space.keys()
    assert isinstance(res["rank"], dict) and res["rank"].keys() == model._search_space["rank"].keys()
    if borders:                              # untouched parameters are pinned to their current value (base_rec.py:207-219)
        assert res["rank"]["args"] == [5, 9] and res["beta"]["args"] == [0.5, 0.5] and res["kind"]["args"] == ["a"]
        assert model._search_space["rank"]["args"] == [8, 256]          # the class attribute is not edited in place


@pytest.mark.parametrize("borders,answer", [(None, True), ({"rank": [-10, -1]}, False)])
def test_param_in_borders(model, borders, answer):   # test_optuna.py:51-56
    assert model._init_params_in_search_space(model._prepare_param_borders(borders)) == answer


def test_it_works(model):   # test_optuna.py:59-67
    assert model._params_tried() is False
    res = model.optimize(LOG, LOG, k=2, budget=1)
    assert isinstance(res["rank"], int) and res == {"rank": 10, "beta": 0.5, "kind": "a"}   # the initial point goes first
    assert model._params_tried() is True
    model.optimize(LOG, LOG, k=2, budget=1)
    assert len(model.study.trials) == 1
    model.optimize(LOG, LOG, k=2, budget=1, new_study=False)
    assert len(model.study.trials) == 2
    assert model.fits == 3


def test_search_stays_inside_borders_and_sets_the_best(model):
    seen = []

    def criterion(recs, test, k):
        seen.append((model.rank, model.beta, model.kind))
        return -abs(model.rank - 40) - model.beta                          # best: rank near 40, small beta
    best = model.optimize(LOG, LOG, param_borders={"rank": [16, 64], "beta": [0.25, 0.75]}, criterion=criterion,
                          k=2, budget=12)
    assert len(seen) == 12 and len(model.study.trials) == 12
    for rank, beta, kind in seen:            # rank / beta searched inside the borders, kind pinned
        assert isinstance(rank, int) and 16 <= rank <= 64 and 0.25 <= beta <= 0.75 and kind == "a"
    assert len({r for r, _, _ in seen}) > 3
    top = max(model.study.trials, key=lambda t: t.value)
    assert best == top.params == {"rank": model.rank, "beta": model.beta, "kind": model.kind}
    assert model.study.best_value == max(-abs(r - 40) - b for r, b, _ in seen)


def test_no_search_space_returns_none(caplog):   # base_rec.py:109-113
    assert PopLike().optimize(LOG, LOG, k=2, budget=1) is None


def test_study_is_picklable_for_save(model):   # model_handler.py:53 dumps `study` with joblib
    model.optimize(LOG, LOG, k=2, budget=2)
    again = pickle.loads(pickle.dumps(model.study))
    assert isinstance(again, RandomSearchStudy) and again.best_params == model.study.best_params


@pytest.mark.parametrize("seed", [0, 1])
def test_default_criterion_is_the_reference_ndcg(seed):
    """`ndcg_at_k` (the default criterion) against the metrics oracle, which is pinned by the reference's golden values."""
    rng = np.random.default_rng(seed)
    recs = pd.DataFrame({"user_idx": rng.integers(0, 30, 400), "item_idx": rng.integers(0, 50, 400),
                         "relevance": rng.random(400)}).drop_duplicates(["user_idx", "item_idx"])
    test = pd.DataFrame({"user_idx": rng.integers(0, 40, 200), "item_idx": rng.integers(0, 50, 200)}).drop_duplicates()
    for k in (1, 5, 10):
        top = recs.sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True])
        by_user = {u: g["item_idx"].tolist()[:k] for u, g in top.groupby("user_idx")}
        truth = {u: g["item_idx"].tolist() for u, g in test.groupby("user_idx")}
        assert ndcg_at_k(recs, test, k) == pytest.approx(metrics_oracle.rank_metrics(by_user, truth, [k])["NDCG"][k], abs=1e-12)
    assert ndcg_at_k(recs.iloc[:0], test, 3) == 0.0


def test_cql_search_space_plumbing_without_gpu():
    from replay_cql_b200.models import CQL
    cql = CQL()
    space = cql._prepare_param_borders({"gamma": [0.95, 0.99], "n_critics": [2, 3]})
    assert space.keys() == CQL._search_space.keys()
    assert space["gamma"]["args"] == [0.95, 0.99] and space["actor_learning_rate"]["args"] == [cql.actor_learning_rate] * 2
    assert cql._init_params_in_search_space(cql._prepare_param_borders(None)) is True
    assert set(CQL._search_space) <= set(cql._init_args)                 # every searchable name is a constructor argument
    with pytest.raises(ValueError):
        cql._prepare_param_borders({"hidden": [128, 256]})
