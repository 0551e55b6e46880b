"""CPU: the pandas mirror of RePlay's Recommender template behaves like the reference's
(re-targeted from /root/reference/tests/models/test_base_rec.py and test_all_models.py),
and its post-processing equals the line-by-line oracle restatement (oracle/recs_oracle.py)."""
from datetime import datetime

import numpy as np
import pandas as pd
import pytest

from oracle import recs_oracle
from replay_cql_b200.recommender import Recommender, get_top_k_recs

LOG = pd.DataFrame([  # reference fixture `log`, tests/utils.py:59-76
    [0, 0, datetime(2019, 8, 22), 4.0], [0, 2, datetime(2019, 8, 23), 3.0], [0, 1, datetime(2019, 8, 27), 2.0],
    [1, 3, datetime(2019, 8, 24), 3.0], [1, 0, datetime(2019, 8, 25), 4.0], [2, 1, datetime(2019, 8, 26), 5.0],
    [2, 0, datetime(2019, 8, 26), 5.0], [2, 2, datetime(2019, 8, 26), 3.0], [3, 1, datetime(2019, 8, 26), 5.0],
    [3, 0, datetime(2019, 8, 26), 5.0], [3, 0, datetime(2019, 8, 26), 1.0]],
    columns=["user_idx", "item_idx", "timestamp", "relevance"])


class DerivedRec(Recommender):
    """Minimal subclass, as in tests/models/test_base_rec.py:12-35."""

    @property
    def _init_args(self):
        return {}

    def _fit(self, log, user_features=None, item_features=None):
        pass

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        pass


class PopLike(Recommender):
    """Returns ALL user x item pairs with popularity as relevance (Appendix-B style _predict)."""

    @property
    def _init_args(self):
        return {}

    def _fit(self, log, user_features=None, item_features=None):
        self.pop = log.groupby("item_idx")["user_idx"].nunique() / log["user_idx"].nunique()

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        u = users["user_idx"].to_numpy()
        i = items["item_idx"].to_numpy()
        return pd.DataFrame({"user_idx": np.repeat(u, len(i)), "item_idx": np.tile(i, len(u)),
                             "relevance": np.tile(self.pop.reindex(i).fillna(0.0).to_numpy(), len(u))})


@pytest.mark.parametrize("array", [None, [1, 2, 2, 3]])
def test_extract_if_needed(array):   # test_base_rec.py:43-48
    from replay_cql_b200.frames import get_ids
    log = pd.DataFrame({"test": [1, 2, 3]})
    assert sorted(get_ids(array if array is not None else log, "test")["test"]) == [1, 2, 3]
    with pytest.raises(ValueError):
        get_ids(5, "test")


def test_users_items_count_and_str():   # test_base_rec.py:51-66
    model = DerivedRec()
    with pytest.raises(AttributeError):
        model._user_dim
    with pytest.raises(AttributeError):
        model._item_dim
    model.fit(LOG)
    assert model._user_dim == 4 and model._item_dim == 4
    assert model.users_count == 4 and model.items_count == 4
    assert str(model) == "DerivedRec"


def test_filter_seen():   # test_all_models.py:265-278
    model = PopLike()
    model.fit(LOG[LOG["user_idx"] != 0])
    assert len(model.predict(log=LOG, users=[3], k=5)) == 2          # user 3 saw items 0,1 -> 2 left
    assert len(model.predict(log=LOG, users=[0], k=5)) == 0          # cold user is dropped
    model.can_predict_cold_users = True
    assert len(model.predict(log=LOG, users=[0], k=5)) == 1          # saw 0,1,2 -> item 3 left
    assert len(model.predict(log=LOG, users=[0], k=5, filter_seen_items=False)) == 4


def test_predict_pairs_k_and_columns():   # test_all_models.py:128-144
    model = PopLike()
    model.fit(LOG)
    pairs = LOG[["user_idx", "item_idx"]]
    assert model.predict_pairs(pairs, log=LOG, k=2).groupby("user_idx").size().max() <= 2
    assert model.predict_pairs(pairs, log=LOG).groupby("user_idx").size().max() > 2
    with pytest.raises(ValueError, match="pairs must be a dataframe with .*"):     # test_all_models.py:170-174
        model.predict_pairs(LOG[["user_idx", "item_idx", "relevance"]])


def test_recs_file_path(tmp_path):   # test_all_models.py:406-447
    model = PopLike()
    model.fit(LOG)
    direct = model.predict(LOG, k=2)
    path = str(tmp_path / "recs.parquet")
    assert model.predict(LOG, k=2, recs_file_path=path) is None
    pd.testing.assert_frame_equal(direct, pd.read_parquet(path))
    assert list(direct.columns) == ["user_idx", "item_idx", "relevance"]


def test_pyarrow_in_pyarrow_out():
    import pyarrow as pa
    model = PopLike()
    table = pa.Table.from_pandas(LOG)
    model.fit(table)
    out = model.predict(table, k=1)
    assert isinstance(out, pa.Table) and out.column_names == ["user_idx", "item_idx", "relevance"]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_wrapper_equals_oracle_restatement(seed):
    rng = np.random.default_rng(seed)
    U, I, k = 9, 30, 4
    log = pd.DataFrame({"user_idx": rng.integers(0, U - 1, 80), "item_idx": rng.integers(0, I, 80),
                        "timestamp": np.arange(80), "relevance": 1.0})

    class Rand(PopLike):
        def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
            df = super()._predict(log, k, users, items)
            table = np.random.default_rng(seed).standard_normal((U, I))   # relevance = f(user, item)
            df["relevance"] = table[df["user_idx"].to_numpy(), df["item_idx"].to_numpy()]
            return df

    model = Rand()
    model.fit(log)
    users = pd.DataFrame({"user_idx": np.sort(log["user_idx"].unique())})
    items = pd.DataFrame({"item_idx": np.sort(log["item_idx"].unique())})
    allp = model._predict(log, k, users, items)
    ref = recs_oracle.predict_wrap(allp, log, k, users)
    got = model.predict(log, k).sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True]).reset_index(drop=True)
    pd.testing.assert_frame_equal(ref.reset_index(drop=True), got, check_dtype=False)


def test_top_k_doctest_case():   # utils.py:69-92
    log = pd.DataFrame({"user_idx": [1, 1, 1, 2], "item_idx": [2, 3, 4, 1], "relevance": [1.0, 1.0, 0.5, 1.0]})
    out = get_top_k_recs(log, 1)
    assert sorted(out["user_idx"]) == [1, 2] and out[out.user_idx == 2]["item_idx"].iloc[0] == 1


def test_cql_init_args_roundtrip_json():
    import json
    from inspect import getfullargspec
    from replay_cql_b200.models import CQL
    m = CQL(top_k=3, n_epochs=2, batch_size=64, n_critics=3)
    args = json.loads(json.dumps(m._init_args))
    names = getfullargspec(CQL.__init__).args[1:]
    assert sorted(args) == sorted(names)            # model_handler.py:70-80 can rebuild the model
    m2 = CQL(**args)
    assert m2._init_args == m._init_args and str(m2) == "CQL"
    with pytest.raises(ValueError):
        CQL(use_gpu=False)
    with pytest.raises(ValueError):
        CQL(soft_q_backup=True)


class PopSaved(PopLike):
    """PopLike with the persistence hooks of a real model: one constructor argument, one data frame
    (the reference's PopRec keeps `item_popularity` in `_dataframes`, replay/models/pop_rec.py)."""

    def __init__(self, alpha: float = 0.0):
        self.alpha = alpha

    @property
    def _init_args(self):
        return {"alpha": self.alpha}

    @property
    def _dataframes(self):
        return {"item_popularity": self.item_popularity}

    def _fit(self, log, user_features=None, item_features=None):
        super()._fit(log)
        self.item_popularity = (self.pop + self.alpha).rename("relevance").reset_index()

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        self.pop = self.item_popularity.set_index("item_idx")["relevance"]
        return super()._predict(log, k, users, items)

    def _save_model(self, path):
        with open(path, "w") as f:
            f.write("opaque")

    def _load_model(self, path):
        with open(path) as f:
            assert f.read() == "opaque"


def test_save_load_layout_and_equal_preds(tmp_path, monkeypatch):   # test_save_load_models.py:48-70, model_handler.py:29-92
    import json
    import os

    from replay_cql_b200 import model_handler, models
    monkeypatch.setattr(models, "PopSaved", PopSaved, raising=False)       # load() resolves the class by name (:69)
    model = PopSaved(alpha=0.25)
    model.fit(LOG)
    base = model.predict(LOG, 2)
    path = str(tmp_path / "test")
    os.makedirs(path)
    open(os.path.join(path, "stale"), "w").close()                         # prepare_dir wipes an existing directory (:19-26)
    model_handler.save(model, path)
    assert sorted(os.listdir(path)) == ["dataframes", "init_args.json", "model", "study"]
    assert sorted(os.listdir(os.path.join(path, "dataframes"))) == ["fit_items", "fit_users", "item_popularity"]
    with open(os.path.join(path, "init_args.json")) as f:
        assert json.load(f) == {"alpha": 0.25, "_model_name": "PopSaved"}
    loaded = model_handler.load(path)
    assert isinstance(loaded, PopSaved) and loaded.alpha == 0.25 and loaded.study is None
    assert loaded.users_count == model.users_count and loaded.items_count == model.items_count
    pd.testing.assert_frame_equal(loaded.predict(LOG, 2).reset_index(drop=True), base.reset_index(drop=True))


LONG_LOG = pd.DataFrame([  # reference fixture `long_log_with_features`, tests/utils.py:77-98
    [0, 0, datetime(2019, 1, 1), 1.0], [0, 3, datetime(2019, 1, 5), 3.0], [0, 1, datetime(2019, 1, 1), 2.0],
    [0, 4, datetime(2019, 1, 1), 4.0], [1, 0, datetime(2020, 1, 5), 4.0], [1, 2, datetime(2018, 1, 1), 2.0],
    [1, 6, datetime(2019, 1, 1), 4.0], [1, 7, datetime(2020, 1, 1), 4.0], [2, 8, datetime(2019, 1, 1), 3.0],
    [2, 1, datetime(2019, 1, 1), 2.0], [2, 5, datetime(2020, 3, 1), 1.0], [2, 6, datetime(2019, 1, 1), 5.0]],
    columns=["user_idx", "item_idx", "timestamp", "relevance"])


def _fit_predict_selected(model, train_log, inf_log, users):   # test_all_models.py:281-286
    model.fit(train_log)
    return model.predict(log=inf_log, users=users, k=1)


def test_predict_new_users():   # test_all_models.py:289-317: a user absent from training but present in the predict log
    model = PopLike()
    model.can_predict_cold_users = True
    pred = _fit_predict_selected(model, LONG_LOG[LONG_LOG["user_idx"] != 0], LONG_LOG, [0])
    assert len(pred) == 1 and pred["user_idx"].iloc[0] == 0


def test_predict_cold_users():   # test_all_models.py:320-345: the user appears in neither log
    model = PopLike()
    model.can_predict_cold_users = True
    warm = LONG_LOG[LONG_LOG["user_idx"] != 0]
    pred = _fit_predict_selected(model, warm, warm, [0])
    assert len(pred) == 1 and pred["user_idx"].iloc[0] == 0


def test_predict_cold_and_new_filter_out():   # test_all_models.py:348-390: models that cannot score cold users drop them
    model = PopLike()
    assert model.can_predict_cold_users is False
    pred = _fit_predict_selected(model, LONG_LOG[LONG_LOG["user_idx"] != 0], LONG_LOG, [0, 3])
    assert len(pred) == 0
    model.can_predict_cold_users = True
    assert 1 <= len(_fit_predict_selected(model, LONG_LOG[LONG_LOG["user_idx"] != 0], LONG_LOG, [0, 3])) <= 2


def test_predict_and_predict_pairs_to_file(tmp_path):   # test_all_models.py:393-447
    model = PopLike()
    path = str(tmp_path / "pred.parquet")
    assert model.fit_predict(LONG_LOG, k=10, recs_file_path=path) is None
    pd.testing.assert_frame_equal(model.predict(LONG_LOG, k=10).reset_index(drop=True), pd.read_parquet(path))
    pairs = LONG_LOG.loc[LONG_LOG["user_idx"] == 1, ["user_idx", "item_idx"]]
    path2 = str(tmp_path / "pairs.parquet")
    assert model.predict_pairs(pairs=pairs, log=LONG_LOG, recs_file_path=path2) is None
    pd.testing.assert_frame_equal(model.predict_pairs(pairs=pairs, log=LONG_LOG).reset_index(drop=True), pd.read_parquet(path2))


class FilteredTopK(PopLike):
    """A model whose `_predict` itself returns at most k unseen items per user, best first (what CQL's fused scorer
    does) and says so: the template must return the same frame as the generic path computes from ALL pairs."""
    _predict_filters_seen = True

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        full = super()._predict(log, k, users, items)
        if filter_seen_items and log is not None:
            seen = set(zip(log["user_idx"], log["item_idx"]))
            full = full[[(u, i) not in seen for u, i in zip(full["user_idx"], full["item_idx"])]]
        return get_top_k_recs(full, k).reset_index(drop=True)


@pytest.mark.parametrize("filter_seen", [True, False])
@pytest.mark.parametrize("k", [1, 2, 5])
def test_model_that_filters_itself_equals_generic_path(k, filter_seen):
    a, b = PopLike(), FilteredTopK()
    a.fit(LOG)
    b.fit(LOG)
    ra = a.predict(LOG, k=k, filter_seen_items=filter_seen)
    calls = []
    b._filter_seen = lambda **kw: calls.append(1) or kw["recs"]          # must not be needed
    rb = b.predict(LOG, k=k, filter_seen_items=filter_seen)
    pd.testing.assert_frame_equal(ra, rb)
    assert not calls
