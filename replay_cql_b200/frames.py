"""Frame adapters: pandas / pyarrow / (optionally) pyspark -> pandas and back.

The reference moves data between the JVM and Python with ``toPandas()`` over
Arrow (``replay/session_handler.py:47``; pattern ``replay/models/neuromf.py:332``).
pyspark/JVM do not exist in this image, so Spark frames are accepted only if
pyspark is importable; the adapter code for them is written to the documented
contract and is exercised nowhere here (DESIGN.md "out of scope").
"""
from __future__ import annotations

import collections.abc
from typing import Any, Iterable, Optional

import numpy as np
import pandas as pd

try:  # optional
    import pyarrow as pa
except Exception:  # pragma: no cover
    pa = None

try:  # optional, absent in this image
    from pyspark.sql import DataFrame as SparkDataFrame  # type: ignore
except Exception:
    SparkDataFrame = None


def is_spark(df: Any) -> bool:
    return SparkDataFrame is not None and isinstance(df, SparkDataFrame)


def to_pandas(df: Any) -> Optional[pd.DataFrame]:
    """Any supported frame -> pandas (no copy for pandas input)."""
    if df is None:
        return None
    if isinstance(df, pd.DataFrame):
        return df
    if pa is not None and isinstance(df, pa.Table):
        return df.to_pandas()
    if is_spark(df):  # pragma: no cover - needs a JVM
        return df.toPandas()
    raise ValueError(f"Wrong type {type(df)}")


def like_input(result: pd.DataFrame, template: Any) -> Any:
    """Return ``result`` in the flavour of ``template`` (Spark in -> Spark out)."""
    if is_spark(template):  # pragma: no cover - needs a JVM
        return template.sparkSession.createDataFrame(result)
    if pa is not None and isinstance(template, pa.Table):
        return pa.Table.from_pandas(result, preserve_index=False)
    return result


def get_ids(data: Any, column: str) -> pd.DataFrame:
    """Unique ids of ``column`` as a one-column frame (``base_rec.py:541-558`` semantics)."""
    if isinstance(data, pd.DataFrame) or (pa is not None and isinstance(data, pa.Table)) or is_spark(data):
        pdf = to_pandas(data)
        return pd.DataFrame({column: pd.unique(pdf[column])})
    if isinstance(data, collections.abc.Iterable):
        return pd.DataFrame({column: pd.unique(pd.Series(list(data)))})
    raise ValueError(f"Wrong type {type(data)}")


def timestamps_to_int64(col: pd.Series) -> np.ndarray:
    """Timestamps (datetime64 / numeric) -> int64 preserving order."""
    if np.issubdtype(col.dtype, np.datetime64):
        return col.to_numpy().astype("datetime64[ns]").astype(np.int64)
    if col.dtype == object:
        return pd.to_datetime(col).to_numpy().astype("datetime64[ns]").astype(np.int64)
    arr = col.to_numpy()
    if np.issubdtype(arr.dtype, np.integer):
        return arr.astype(np.int64)
    # float timestamps: order-preserving rank
    return np.argsort(np.argsort(arr, kind="stable"), kind="stable").astype(np.int64)
