"""MDP builder: interaction log -> episode-ordered (observation, action, reward, terminal).

Replaces the RePlay CQL wrapper's ``MdpDatasetBuilder.build`` ([EXT] upstream
``replay/experimental/models/cql.py``) -- two Spark window
sorts, a global ``orderBy`` and ``toPandas()`` -- with vectorised sorts over the
Arrow/pandas columns, written straight into (pinned, when CUDA is up) host
buffers that ``cql_load_transitions`` uploads and expands on the GPU.

Semantics (one episode per user):
* global order (user_idx, timestamp) ascending; ties keep input order;
* reward = 1 for the user's ``top_k`` rows by (relevance desc, timestamp desc), else 0;
* terminal = 1 on the user's last row in that order;
* action = relevance + N(0,1) * ``action_randomization_scale``;
* observation = (user_idx, item_idx) as float32 (exact below 2**24).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from .frames import to_pandas, timestamps_to_int64


@dataclass
class MdpArrays:
    obs: np.ndarray    # [n, 2] float32
    act: np.ndarray    # [n] float32
    rew: np.ndarray    # [n] float32
    term: np.ndarray   # [n] float32
    order: np.ndarray  # [n] int64: original row of each step

    def __len__(self) -> int:
        return int(self.obs.shape[0])


def alloc_host(shape, dtype=np.float32) -> np.ndarray:
    """Host buffer for H2D staging: pinned when a CUDA context can be created, else pageable."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy()
    except Exception:  # pragma: no cover
        pass
    return np.empty(shape, dtype=dtype)


def build_mdp(log: Any, top_k: int = 10, action_randomization_scale: float = 1e-3,
              seed: Optional[int] = None, action_noise: Optional[np.ndarray] = None) -> MdpArrays:
    """``action_noise`` (per ORIGINAL row, already scaled) overrides the seeded draw (parity tests)."""
    pdf = to_pandas(log)
    for col in ("user_idx", "item_idx", "timestamp", "relevance"):
        if col not in pdf.columns:
            raise ValueError(f"log must have column {col}")
    n = len(pdf)
    user = pdf["user_idx"].to_numpy().astype(np.int64)
    item = pdf["item_idx"].to_numpy().astype(np.int64)
    if n and (user.max() >= 2 ** 24 or item.max() >= 2 ** 24):
        raise ValueError("user_idx/item_idx must be < 2**24 to be exact in float32 observations")
    ts = timestamps_to_int64(pdf["timestamp"])
    rel = pdf["relevance"].to_numpy().astype(np.float64)
    row = np.arange(n, dtype=np.int64)

    order = np.lexsort((row, ts, user))                      # (user, timestamp, input order)
    u_sorted = user[order]
    term = np.ones(n, dtype=np.float32)
    if n > 1:
        term[:-1] = (u_sorted[1:] != u_sorted[:-1]).astype(np.float32)

    rank_order = np.lexsort((row, -ts, -rel, user))          # (user, relevance desc, timestamp desc)
    ur = user[rank_order]
    starts = np.flatnonzero(np.r_[True, ur[1:] != ur[:-1]]) if n else np.zeros(0, dtype=np.int64)
    group_start = np.repeat(starts, np.diff(np.r_[starts, n])) if n else starts
    pos_in_user = np.arange(n, dtype=np.int64) - group_start
    rewarded = np.zeros(n, dtype=np.float32)
    rewarded[rank_order] = (pos_in_user < top_k).astype(np.float32)

    if action_noise is None:
        rng = np.random.default_rng(seed)
        action_noise = rng.standard_normal(n) * action_randomization_scale
    action = rel.astype(np.float32).astype(np.float64) + np.asarray(action_noise, dtype=np.float64)

    obs = alloc_host((n, 2))
    obs[:, 0] = u_sorted
    obs[:, 1] = item[order]
    act = alloc_host((n,)); act[:] = action[order]
    rew = alloc_host((n,)); rew[:] = rewarded[order]
    trm = alloc_host((n,)); trm[:] = term
    return MdpArrays(obs, act, rew, trm, order)


def to_transitions(mdp: MdpArrays) -> dict:
    """Host-side expansion into (s, a, r, s', done) -- what the GPU loader builds; for tests / e2e batches."""
    n = len(mdp)
    nxt = np.zeros((n, 2), dtype=np.float32)
    if n > 1:
        nxt[:-1] = mdp.obs[1:]
    nxt[mdp.term > 0] = 0.0
    return {"obs": mdp.obs, "act": mdp.act.reshape(-1, 1), "rew": mdp.rew.reshape(-1, 1),
            "next_obs": nxt, "term": mdp.term.reshape(-1, 1)}


def seen_csr(log: Any, n_users_dim: int):
    """Per-user sorted, de-duplicated seen items as CSR over user id (for the lazy seen filter)."""
    pdf = to_pandas(log)
    user = pdf["user_idx"].to_numpy().astype(np.int64)
    item = pdf["item_idx"].to_numpy().astype(np.int64)
    if user.size:
        n_users_dim = max(n_users_dim, int(user.max()) + 1)
    key = np.unique(user * (2 ** 32) + item)
    u, i = key >> 32, (key & 0xFFFFFFFF).astype(np.int32)
    counts = np.bincount(u, minlength=n_users_dim)
    indptr = np.zeros(n_users_dim + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, i


def build_mdp_on_device(engine, log: Any, top_k: int = 10, action_randomization_scale: float = 1e-3,
                        action_noise: Optional[np.ndarray] = None, want_outputs: bool = False):
    """Same semantics as :func:`build_mdp`, but sorted and expanded on the GPU (``cql_build_mdp``): the log
    columns go host -> device once and the replay table never exists on the host."""
    pdf = to_pandas(log)
    for col in ("user_idx", "item_idx", "timestamp", "relevance"):
        if col not in pdf.columns:
            raise ValueError(f"log must have column {col}")
    return engine.build_mdp_on_device(pdf["user_idx"].to_numpy(), pdf["item_idx"].to_numpy(),
                                      timestamps_to_int64(pdf["timestamp"]), pdf["relevance"].to_numpy(),
                                      top_k=top_k, action_randomization_scale=action_randomization_scale,
                                      action_noise=action_noise, want_outputs=want_outputs)
