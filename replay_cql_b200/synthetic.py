"""Seeded synthetic interaction logs of the BASELINE.json shapes.

Users: log-normal history length (min 20, like ML-1M); items: Zipf(1.0) popularity;
ratings in {1..5} with an ML-like pmf; timestamps strictly increasing per user; indices
dense from 0 (Indexer semantics, ``replay/data_preparator.py:77-82``).  Generator seed
12345 matches ``SEED`` of the reference's model-comparison notebook.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

SHAPES = {
    "ml1m": dict(n_users=6040, n_items=3706, n_rows=1_000_209),
    "ml20m": dict(n_users=138_493, n_items=26_744, n_rows=20_000_263),
    "tiny": dict(n_users=64, n_items=300, n_rows=4_000),
    # BASELINE.json configs[4]; too large for a host-side generator -- CqlEngine.synth_table builds it in HBM
    "stress": dict(n_users=10_000_000, n_items=1_000_000, n_rows=1_000_000_000),
}
RATING_PMF = np.array([0.06, 0.11, 0.26, 0.35, 0.22])


def history_lengths(n_users: int, n_rows: int, rng: np.random.Generator, min_len: int = 20) -> np.ndarray:
    min_len = min(min_len, max(1, n_rows // n_users))
    raw = rng.lognormal(mean=0.0, sigma=1.0, size=n_users)
    spare = n_rows - min_len * n_users
    extra = np.floor(raw / raw.sum() * spare).astype(np.int64)
    lens = min_len + extra
    deficit = n_rows - int(lens.sum())
    lens[rng.choice(n_users, size=deficit, replace=True)] += 1 if deficit > 0 else 0
    # choice with replacement may hit a user twice: fix the remainder deterministically
    rem = n_rows - int(lens.sum())
    i = 0
    while rem > 0:
        lens[i % n_users] += 1
        rem -= 1
        i += 1
    return lens


def make_log_arrays(shape: str = "ml1m", seed: int = 12345, n_rows: int | None = None):
    """-> dict(user_idx int32, item_idx int32, timestamp int64, relevance float64), user-major order."""
    cfg = dict(SHAPES[shape])
    if n_rows is not None:
        cfg["n_rows"] = n_rows
    rng = np.random.default_rng(seed)
    U, I, N = cfg["n_users"], cfg["n_items"], cfg["n_rows"]
    lens = history_lengths(U, N, rng)
    user = np.repeat(np.arange(U, dtype=np.int32), lens)
    pop = 1.0 / np.arange(1, I + 1, dtype=np.float64)
    cdf = np.cumsum(pop / pop.sum())
    item = np.searchsorted(cdf, rng.random(N), side="left").astype(np.int32)
    np.minimum(item, I - 1, out=item)
    rating = (np.searchsorted(np.cumsum(RATING_PMF), rng.random(N), side="left") + 1).astype(np.float64)
    starts = np.cumsum(lens) - lens
    ts = (np.arange(N, dtype=np.int64) - np.repeat(starts, lens)) * 60 + 946_684_800  # +1 minute per event
    return {"user_idx": user, "item_idx": item, "timestamp": ts, "relevance": rating,
            "n_users": U, "n_items": I}


def make_log(shape: str = "ml1m", seed: int = 12345, n_rows: int | None = None) -> pd.DataFrame:
    a = make_log_arrays(shape, seed, n_rows)
    return pd.DataFrame({k: a[k] for k in ("user_idx", "item_idx", "timestamp", "relevance")})
