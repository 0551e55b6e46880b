"""Spark-facing ``CQL`` for a real RePlay installation (``from replay_cql_b200.spark import CQL``).

Written to the documented contract of ``replay/models/base_rec.py`` (``Recommender``: ``_fit`` :376,
``_predict`` :607, ``_init_args`` :143-149, ``_save_model``/``_load_model`` :280-284) and NEVER executed in this
image: pyspark and a JVM are absent.  It reuses the engine wiring of ``models.CQL``
and only converts frames at the boundary: ``toPandas()`` in (Arrow, ``replay/session_handler.py:47``),
``createDataFrame(..., REC_SCHEMA)`` out (``replay/constants.py:25-31``).
"""
from __future__ import annotations

try:  # pragma: no cover - needs pyspark + replay
    from replay.constants import REC_SCHEMA
    from replay.models.base_rec import Recommender as _SparkRecommender
    from replay.session_handler import State
except Exception as exc:  # pragma: no cover
    raise ImportError("replay_cql_b200.spark needs an importable RePlay (pyspark + JVM); "
                      "use replay_cql_b200.models.CQL on pandas/pyarrow frames instead") from exc

from .models import CQL as _PandasCQL  # pragma: no cover


class CQL(_SparkRecommender):  # pragma: no cover
    """Same constructor and hooks as ``models.CQL``; Spark DataFrames in and out."""

    _search_space = _PandasCQL._search_space

    def __init__(self, *args, **kwargs):
        self._impl = _PandasCQL(*args, **kwargs)

    @property
    def _init_args(self):
        return self._impl._init_args

    def __getattr__(self, name):
        # hyper-parameters live on the pandas model (only called when normal lookup fails)
        impl = self.__dict__.get("_impl")
        if impl is not None and name in impl._init_args:
            return getattr(impl, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        # `set_params` (base_rec.py:151-163) does setattr on the wrapper: optimize() trials must reach the engine's model
        impl = self.__dict__.get("_impl")
        if impl is not None and name in impl._init_args:
            setattr(impl, name, value)
        else:
            object.__setattr__(self, name, value)

    def _fit(self, log, user_features=None, item_features=None) -> None:
        self._impl._fit(log.select("user_idx", "item_idx", "timestamp", "relevance").toPandas())

    def _predict(self, log, k, users, items, user_features=None, item_features=None, filter_seen_items=True):
        recs = self._impl._predict(log.toPandas() if log is not None else None, k, users.toPandas(),
                                   items.toPandas(), None, None, filter_seen_items)
        return State().session.createDataFrame(recs, schema=REC_SCHEMA)

    def _predict_pairs(self, pairs, log=None, user_features=None, item_features=None):
        recs = self._impl._predict_pairs(pairs.toPandas(), None)
        return State().session.createDataFrame(recs, schema=REC_SCHEMA)

    def _save_model(self, path: str) -> None:
        self._impl._save_model(path)

    def _load_model(self, path: str) -> None:
        self._impl._load_model(path)
