"""pandas/pyarrow mirror of RePlay's ``Recommender`` template methods.

The reference base class (``replay/models/base_rec.py``) is written against
PySpark, which cannot run in this image (no JVM).  This module restates the
template -- same method names, argument meaning, error behaviour and result
schema -- on pandas, so a model written against it (``models.CQL``) is a
drop-in for the hot path, and so the reference's conformance tests can be
re-targeted here (``tests/test_recommender_conformance.py``).  Reference lines
are cited per method.  Frames may be pandas, pyarrow Tables or (if pyspark is
importable) Spark DataFrames; results come back in the flavour of the input.
"""
from __future__ import annotations

import logging
import math
from abc import ABC, abstractmethod
from copy import deepcopy
from typing import Any, Callable, Dict, Iterable, List, Optional, Union

import numpy as np
import pandas as pd

from .frames import get_ids, like_input, to_pandas

REC_COLUMNS = ["user_idx", "item_idx", "relevance"]  # replay/constants.py:25-31 (REC_SCHEMA)


def _rec_frame(users, items, relevance) -> pd.DataFrame:
    return pd.DataFrame({
        "user_idx": np.asarray(users, dtype=np.int32),
        "item_idx": np.asarray(items, dtype=np.int32),
        "relevance": np.asarray(relevance, dtype=np.float64),
    })


def get_top_k(df: pd.DataFrame, partition_by: str, order_by: list, ascending: list, k: int) -> pd.DataFrame:
    """``replay/utils.py:59-109``: ``row_number() over (partition ... order by ...) <= k``."""
    if len(df) == 0:
        return df
    ordered = df.sort_values([partition_by] + order_by, ascending=[True] + ascending, kind="stable")
    rank = ordered.groupby(partition_by, sort=False).cumcount()
    return ordered[rank < k]


def get_top_k_recs(recs: pd.DataFrame, k: int) -> pd.DataFrame:
    """``replay/utils.py:112-127``.  Spark leaves ties arbitrary; here ties go to the lower item_idx."""
    return get_top_k(recs, "user_idx", ["relevance", "item_idx"], [False, True], k)


def ndcg_at_k(recs: pd.DataFrame, ground_truth: pd.DataFrame, k: int) -> float:
    """Default ``optimize`` criterion: the reference's ``NDCG()(recs, test, k)`` on pandas -- per-user body of
    ``replay/metrics/ndcg.py:51-61``, users = the ground-truth users with missing predictions counted as zero
    (``replay/metrics/base_metric.py:102-140``), ``k`` best predictions by ``get_top_k_recs``."""
    gt = ground_truth.groupby("user_idx")["item_idx"].agg(lambda x: set(x.tolist()))
    if len(gt) == 0:
        return 0.0
    top = get_top_k_recs(recs, k) if len(recs) else recs
    top = top.sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True], kind="stable")
    pred = top.groupby("user_idx")["item_idx"].agg(list) if len(top) else pd.Series(dtype=object)
    total = 0.0
    for user, truth in gt.items():
        items = pred.get(user, [])
        if not items or not truth:
            continue
        dcg = sum(1.0 / math.log2(i + 2) for i, item in enumerate(items[:k]) if item in truth)
        idcg = sum(1.0 / math.log2(i + 2) for i in range(min(k, len(truth))))
        total += dcg / idcg
    return total / len(gt)


class _Trial:
    """What ``suggest_params`` needs from an optuna trial (``replay/optuna_objective.py:47-73``)."""

    def __init__(self, rng: np.random.Generator, fixed: Optional[dict] = None):
        self.rng, self.fixed, self.params, self.value = rng, fixed or {}, {}, None

    def _keep(self, name, value):
        self.params[name] = self.fixed.get(name, value)
        return self.params[name]

    def suggest_uniform(self, name, low, high):
        return self._keep(name, low if low == high else float(self.rng.uniform(low, high)))

    def suggest_loguniform(self, name, low, high):     # a pinned parameter ([v, v]) comes back bit-identical
        return self._keep(name, low if low == high else float(math.exp(self.rng.uniform(math.log(low), math.log(high)))))

    def suggest_int(self, name, low, high, log=False):
        if low == high:
            return self._keep(name, low)
        if log:
            return self._keep(name, int(min(high, max(low, round(math.exp(self.rng.uniform(math.log(low), math.log(high + 1)) - 0.5))))))
        return self._keep(name, int(self.rng.integers(low, high + 1)))

    def suggest_categorical(self, name, choices):
        return self._keep(name, choices[int(self.rng.integers(len(choices)))])


class RandomSearchStudy:
    """Stand-in for ``optuna.create_study(direction="maximize")`` when optuna is not installed: the same members
    ``optimize`` touches (``enqueue_trial``, ``optimize(objective, n_trials)``, ``trials[i].params``, ``best_params``,
    ``best_value``), seeded random sampling instead of TPE.  Picklable, so ``save`` / ``load`` keep it as ``study``."""

    def __init__(self, seed: int = 0):
        self.rng = np.random.default_rng(seed)
        self.trials: List[_Trial] = []
        self._queue: List[dict] = []

    def enqueue_trial(self, params: dict) -> None:
        self._queue.append(dict(params))

    def optimize(self, objective: Callable, n_trials: int) -> None:
        for _ in range(n_trials):
            trial = _Trial(self.rng, self._queue.pop(0) if self._queue else None)
            trial.value = float(objective(trial))
            self.trials.append(trial)

    @property
    def best_trial(self) -> _Trial:
        if not self.trials:
            raise ValueError("No trials are completed yet.")
        return max(self.trials, key=lambda t: t.value)

    @property
    def best_params(self) -> dict:
        return dict(self.best_trial.params)

    @property
    def best_value(self) -> float:
        return self.best_trial.value


def _create_study():
    try:                                      # the reference's study (base_rec.py:116-119) when optuna is there
        from optuna import create_study
        from optuna.samplers import TPESampler
        return create_study(direction="maximize", sampler=TPESampler())
    except ImportError:
        return RandomSearchStudy()


def suggest_params(trial, search_space: Dict[str, Dict[str, Any]]) -> Dict[str, Any]:
    """``replay/optuna_objective.py:47-73``."""
    suggest = {"uniform": trial.suggest_uniform, "int": trial.suggest_int, "loguniform": trial.suggest_loguniform,
               "loguniform_int": lambda name, low, high: trial.suggest_int(name, low, high, log=True)}
    res = {}
    for param, spec in search_space.items():
        if spec["type"] == "categorical":
            res[param] = trial.suggest_categorical(param, spec["args"])
        else:
            low, high = spec["args"]
            res[param] = suggest[spec["type"]](param, low, high)
    return res


class Recommender(ABC):
    """Base class for models that fit on an interaction log (``base_rec.py:59-78, 1202-1335``)."""

    can_predict_cold_users: bool = False
    can_predict_cold_items: bool = False
    _search_space: Optional[Dict[str, Dict[str, Any]]] = None
    _logger: Optional[logging.Logger] = None
    study = None

    # ------------------------------------------------------------ hooks (base_rec.py:143-149, 276-284, 376, 607)
    @property
    @abstractmethod
    def _init_args(self) -> dict:
        """Constructor kwargs, JSON-serialisable (``model_handler.py:40-43``)."""

    @property
    def _dataframes(self) -> dict:
        return {}

    def _save_model(self, path: str) -> None:
        pass

    def _load_model(self, path: str) -> None:
        pass

    @abstractmethod
    def _fit(self, log: pd.DataFrame, user_features=None, item_features=None) -> None:
        ...

    @abstractmethod
    def _predict(self, log: Optional[pd.DataFrame], k: int, users: pd.DataFrame, items: pd.DataFrame,
                 user_features=None, item_features=None, filter_seen_items: bool = True) -> pd.DataFrame:
        ...

    def _clear_cache(self) -> None:
        pass

    # ------------------------------------------------------------ small public surface
    @property
    def logger(self) -> logging.Logger:
        if self._logger is None:
            self._logger = logging.getLogger("replay")
        return self._logger

    def set_params(self, **params: Any) -> None:
        """``base_rec.py:315-324``."""
        for param, value in params.items():
            setattr(self, param, value)
        self._clear_cache()

    # ------------------------------------------------------------ hyper-parameter search (base_rec.py:80-274, 938-952)
    def _default_criterion(self) -> Callable[[pd.DataFrame, pd.DataFrame, int], float]:
        return ndcg_at_k

    def optimize(self, train: Any, test: Any, user_features=None, item_features=None,
                 param_borders: Optional[Dict[str, List[Any]]] = None, criterion: Optional[Callable] = None,
                 k: int = 10, budget: int = 10, new_study: bool = True) -> Optional[Dict[str, Any]]:
        """``base_rec.py:80-141``: search ``_search_space`` (narrowed by ``param_borders``) for the parameters that
        maximise ``criterion(recs, test, k)`` (default NDCG) of ``fit(train)`` + ``predict`` for the test users and
        items; the initial parameters are tried first if they lie inside the space; the best ones are set on the
        model and returned."""
        if self._search_space is None:
            self.logger.warning("%s has no hyper parameters to optimize", str(self))
            return None
        if self.study is None or new_study:
            self.study = _create_study()
        search_space = self._prepare_param_borders(param_borders)
        if self._init_params_in_search_space(search_space) and not self._params_tried():
            self.study.enqueue_trial({p: v for p, v in self._init_args.items() if p in search_space})
        train_pd, test_pd = to_pandas(train), to_pandas(test)
        users = test_pd[["user_idx"]].drop_duplicates()
        items = test_pd[["item_idx"]].drop_duplicates()
        metric = criterion if criterion is not None else self._default_criterion()

        def objective(trial) -> float:            # optuna_objective.py:76-134 (eval_quality, scenario_objective_calculator)
            self.set_params(**suggest_params(trial, search_space))
            self._fit_wrap(train_pd, user_features, item_features)
            recs = self._predict_wrap(log=train_pd, k=k, users=users, items=items,
                                      user_features=user_features, item_features=item_features)
            return float(metric(recs, test_pd, k))

        self.study.optimize(objective, budget)
        best_params = self.study.best_params
        self.set_params(**best_params)
        return best_params

    def _init_params_in_search_space(self, search_space: Dict[str, Dict[str, Any]]) -> bool:
        """``base_rec.py:151-182``."""
        outside = {}
        for param, value in self._init_args.items():
            if param not in search_space:
                continue
            borders, kind = search_space[param]["args"], search_space[param]["type"]
            if (kind == "categorical" and value not in borders) or \
                    (kind != "categorical" and (value < borders[0] or value > borders[1])):
                outside[param] = {"borders": borders, "value": value}
        if outside:
            self.logger.debug("Model is initialized with parameters outside the search space: %s."
                              "Initial parameters will not be evaluated during optimization."
                              "Change search spare with 'param_borders' argument if necessary", outside)
            return False
        return True

    def _prepare_param_borders(self, param_borders: Optional[Dict[str, List[Any]]] = None) -> Dict[str, Dict[str, Any]]:
        """``base_rec.py:184-221``: parameters without user borders are pinned to their current value."""
        search_space = deepcopy(self._search_space)
        if param_borders is None:
            return search_space
        for param, borders in param_borders.items():
            self._check_borders(param, borders)
            search_space[param]["args"] = borders
        args = self._init_args
        for param in search_space:
            if param not in param_borders:
                value = args[param]
                search_space[param]["args"] = [value] if search_space[param]["type"] == "categorical" else [value, value]
        return search_space

    def _check_borders(self, param: str, borders: Any) -> None:
        """``base_rec.py:223-241``."""
        if param not in self._search_space:
            raise ValueError(f"Hyper parameter {param} is not defined for {str(self)}")
        if not isinstance(borders, list):
            raise ValueError(f"Parameter {param} borders are not a list")
        if self._search_space[param]["type"] != "categorical" and len(borders) != 2:
            raise ValueError(f"Hyper parameter {param} is numerical but bounds are not in ([lower, upper]) format")

    def _params_tried(self) -> bool:
        """``base_rec.py:938-952``."""
        if self.study is None:
            return False
        params = {name: value for name, value in self._init_args.items() if name in self._search_space}
        return any(params == trial.params for trial in self.study.trials)

    def __str__(self) -> str:
        return type(self).__name__

    def _get_fit_counts(self, entity: str) -> int:
        if not hasattr(self, f"_num_{entity}s"):
            setattr(self, f"_num_{entity}s", len(getattr(self, f"fit_{entity}s")))
        return getattr(self, f"_num_{entity}s")

    @property
    def users_count(self) -> int:
        return self._get_fit_counts("user")

    @property
    def items_count(self) -> int:
        return self._get_fit_counts("item")

    def _get_fit_dims(self, entity: str) -> int:
        """max idx + 1, recomputed lazily after ``load`` (``base_rec.py:671-695``)."""
        if not hasattr(self, f"_{entity}_dim_size"):
            frame = getattr(self, f"fit_{entity}s")  # AttributeError before fit, like the reference
            setattr(self, f"_{entity}_dim_size", int(frame[f"{entity}_idx"].max()) + 1)
        return getattr(self, f"_{entity}_dim_size")

    @property
    def _user_dim(self) -> int:
        return self._get_fit_dims("user")

    @property
    def _item_dim(self) -> int:
        return self._get_fit_dims("item")

    # ------------------------------------------------------------ fit (base_rec.py:329-373, 1205-1217)
    def fit(self, log: Any) -> None:
        self._fit_wrap(log, None, None)

    def _fit_wrap(self, log: Any, user_features=None, item_features=None) -> None:
        self.logger.debug("Starting fit %s", type(self).__name__)
        from . import frames
        if frames.pa is not None and isinstance(log, frames.pa.Table) and getattr(self, "accepts_arrow", False):
            # Arrow in: distinct ids with pyarrow.compute, the table itself goes to `_fit` -- no `toPandas()` at all
            # (the step being replaced: replay/models/neuromf.py:332)
            import pyarrow.compute as pc
            uniq = lambda c: np.sort(pc.unique(log.column(c)).to_numpy(zero_copy_only=False))
            self.fit_users = pd.DataFrame({"user_idx": uniq("user_idx")})
            self.fit_items = pd.DataFrame({"item_idx": uniq("item_idx")})
            pdf = log
        else:
            pdf = to_pandas(log)
            self.fit_users = pd.DataFrame({"user_idx": np.sort(pd.unique(pdf["user_idx"]))})
            self.fit_items = pd.DataFrame({"item_idx": np.sort(pd.unique(pdf["item_idx"]))})
        self._num_users = len(self.fit_users)
        self._num_items = len(self.fit_items)
        self._user_dim_size = int(self.fit_users["user_idx"].max()) + 1
        self._item_dim_size = int(self.fit_items["item_idx"].max()) + 1
        self._fit(pdf, user_features, item_features)

    # ------------------------------------------------------------ cold filtering (base_rec.py:560-603)
    def _filter_cold(self, df: Optional[pd.DataFrame], entity: str):
        if getattr(self, f"can_predict_cold_{entity}s") or df is None:
            return 0, df
        col = f"{entity}_idx"
        known = getattr(self, f"fit_{entity}s")[col]
        mask = df[col].isin(known)
        num_cold = int(df.loc[~mask, col].nunique())
        if num_cold == 0:
            return 0, df
        return num_cold, df[mask]

    def _filter_cold_for_predict(self, main_df, log_df, entity: str):
        num_new, main_df = self._filter_cold(main_df, entity)
        if num_new > 0:
            self.logger.info("%s model can't predict cold %ss, they will be ignored", self, entity)
        _, log_df = self._filter_cold(log_df, entity)
        return main_df, log_df

    # ------------------------------------------------------------ seen filter (base_rec.py:417-464)
    def _filter_seen(self, recs: pd.DataFrame, log: pd.DataFrame, k: int, users: pd.DataFrame) -> pd.DataFrame:
        """Drop items present in ``log`` from each user's recs, after the reference's crop to k + seen."""
        users_log = log.merge(users[["user_idx"]].drop_duplicates(), on="user_idx")
        num_seen = users_log.groupby("user_idx")["item_idx"].count().rename("seen_count").reset_index()
        max_seen = int(num_seen["seen_count"].max()) if len(num_seen) else 0
        if len(recs) == 0:
            return recs
        ordered = recs.sort_values(["user_idx", "relevance", "item_idx"], ascending=[True, False, True], kind="stable")
        ordered = ordered.assign(temp_rank=ordered.groupby("user_idx", sort=False).cumcount() + 1)
        ordered = ordered[ordered["temp_rank"] <= max_seen + k]
        ordered = ordered.merge(num_seen, on="user_idx", how="left")
        ordered["seen_count"] = ordered["seen_count"].fillna(0)
        ordered = ordered[ordered["temp_rank"] <= ordered["seen_count"] + k].drop(columns=["temp_rank", "seen_count"])
        seen_key = users_log["user_idx"].to_numpy().astype(np.int64) * (2 ** 32) + users_log["item_idx"].to_numpy().astype(np.int64)
        rec_key = ordered["user_idx"].to_numpy().astype(np.int64) * (2 ** 32) + ordered["item_idx"].to_numpy().astype(np.int64)
        return ordered[~np.isin(rec_key, seen_key)]

    # ------------------------------------------------------------ predict (base_rec.py:467-539, 1220-1257)
    def predict(self, log: Any, k: int, users: Optional[Union[Any, Iterable]] = None,
                items: Optional[Union[Any, Iterable]] = None, filter_seen_items: bool = True,
                recs_file_path: Optional[str] = None):
        return self._predict_wrap(log, k, users, items, None, None, filter_seen_items, recs_file_path)

    def _predict_wrap(self, log, k: int, users=None, items=None, user_features=None, item_features=None,
                      filter_seen_items: bool = True, recs_file_path: Optional[str] = None):
        self.logger.debug("Starting predict %s", type(self).__name__)
        template = log
        log_pdf = to_pandas(log)
        user_data = users if users is not None else (log_pdf if log_pdf is not None else self.fit_users)
        users_df = get_ids(user_data, "user_idx")
        users_df, log_pdf = self._filter_cold_for_predict(users_df, log_pdf, "user")
        item_data = items if items is not None else self.fit_items
        items_df = get_ids(item_data, "item_idx")
        items_df, log_pdf = self._filter_cold_for_predict(items_df, log_pdf, "item")
        if len(items_df) < k:
            self.logger.debug("k = %d > number of items = %d", k, len(items_df))
        recs = self._predict(log_pdf, k, users_df, items_df, user_features, item_features, filter_seen_items)
        # (a model whose `_predict` already returns at most k UNSEEN items per user -- CQL's fused scorer -- declares it:
        #  the generic pass would sort the whole log again to drop nothing)
        if filter_seen_items and log_pdf is not None and not getattr(self, "_predict_filters_seen", False):
            recs = self._filter_seen(recs=recs, log=log_pdf, users=users_df, k=k)
        if getattr(self, "_predict_filters_seen", False) and self._is_top_k(recs, k):
            recs = recs[REC_COLUMNS].reset_index(drop=True)       # already the k best per user, best first: nothing to rank
        else:
            recs = get_top_k_recs(recs, k)[REC_COLUMNS].reset_index(drop=True)
        if recs_file_path is not None:
            recs.to_parquet(recs_file_path, index=False)
            return None
        return like_input(recs, template)

    @staticmethod
    def _is_top_k(recs: pd.DataFrame, k: int) -> bool:
        """True when ``recs`` is grouped by ascending user, at most ``k`` rows per user, relevance descending inside a
        user with ties by ascending item -- i.e. ``get_top_k_recs`` would return it unchanged (checked in O(n), no sort)."""
        if len(recs) < 2:
            return True
        u = recs["user_idx"].to_numpy()
        r = recs["relevance"].to_numpy()
        i = recs["item_idx"].to_numpy()
        same = u[1:] == u[:-1]
        if not np.all((u[1:] > u[:-1]) | same):
            return False
        if not np.all(~same | (r[1:] < r[:-1]) | ((r[1:] == r[:-1]) & (i[1:] > i[:-1]))):
            return False
        starts = np.flatnonzero(np.concatenate(([True], ~same)))
        return int(np.diff(np.concatenate((starts, [len(u)]))).max()) <= k

    def fit_predict(self, log: Any, k: int, users=None, items=None, filter_seen_items: bool = True,
                    recs_file_path: Optional[str] = None):
        """``base_rec.py:1287-1323``."""
        self._fit_wrap(log, None, None)
        return self._predict_wrap(log, k, users, items, None, None, filter_seen_items, recs_file_path)

    # ------------------------------------------------------------ predict_pairs (base_rec.py:725-823, 1259-1285)
    def predict_pairs(self, pairs: Any, log: Any = None, recs_file_path: Optional[str] = None,
                      k: Optional[int] = None):
        template = pairs
        pairs_pdf, log_pdf = to_pandas(pairs), to_pandas(log)
        if sorted(pairs_pdf.columns) != ["item_idx", "user_idx"]:
            raise ValueError("pairs must be a dataframe with columns strictly [user_idx, item_idx]")
        pairs_pdf, log_pdf = self._filter_cold_for_predict(pairs_pdf, log_pdf, "user")
        pairs_pdf, log_pdf = self._filter_cold_for_predict(pairs_pdf, log_pdf, "item")
        pred = self._predict_pairs(pairs_pdf, log_pdf)
        if k:
            pred = get_top_k(pred, "user_idx", ["relevance", "item_idx"], [False, True], k)
        pred = pred[REC_COLUMNS].reset_index(drop=True)
        if recs_file_path is not None:
            pred.to_parquet(recs_file_path, index=False)
            return None
        return like_input(pred, template)

    def _predict_pairs(self, pairs: pd.DataFrame, log: Optional[pd.DataFrame] = None) -> pd.DataFrame:
        """Generic fallback (``base_rec.py:784-823``): full predict joined with ``pairs``."""
        self.logger.warning("native predict_pairs is not implemented for this model. "
                            "Falling back to usual predict method and filtering the results.")
        users = pd.DataFrame({"user_idx": pd.unique(pairs["user_idx"])})
        items = pd.DataFrame({"item_idx": pd.unique(pairs["item_idx"])})
        pred = self._predict(log, len(items), users, items, None, None, False)
        return pred.merge(pairs[["user_idx", "item_idx"]], on=["user_idx", "item_idx"], how="inner")
