"""CPU restatement of the RePlay CQL wrapper's MDP builder.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED ([EXT-UNVERIFIED]): the builder is not in
the mounted checkout; this follows the published upstream descendant
(``replay/experimental/models/cql.py::MdpDatasetBuilder.build``) and d3rlpy 1.x
``dataset.pyx::_to_transitions`` for the episode -> transition expansion.
The input schema is the reference's ``LOG_SCHEMA`` (``replay/constants.py:16-23``).

Deliberately written as slow, obvious per-user Python loops over a pandas
frame -- it shares nothing with the vectorised product builder.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def build_mdp(log: pd.DataFrame, top_k: int = 10, action_noise: np.ndarray | None = None):
    """log[user_idx,item_idx,timestamp,relevance] -> dict of float32 arrays.

    * global order: (user_idx, timestamp) ascending, ties keep input order;
    * reward = 1 for the user's ``top_k`` rows by (relevance desc, timestamp desc), else 0;
    * terminal = 1 on the user's last row in that order;
    * action = relevance + action_noise[row]  (noise indexed by ORIGINAL row position;
      the wrapper draws ``randn()*action_randomization_scale`` -- an input here);
    * transitions: next_obs = next row of the same user, zeros on the terminal row.
    """
    df = log.reset_index(drop=True).copy()
    n = len(df)
    if action_noise is None:
        action_noise = np.zeros(n, dtype=np.float64)
    df["row_"] = np.arange(n)
    df["action_"] = df["relevance"].astype(np.float32).astype(np.float64) + action_noise
    obs, act, rew, term, nxt = [], [], [], [], []
    for user, grp in df.groupby("user_idx", sort=True):
        g = grp.sort_values(["timestamp", "row_"], kind="stable")
        rows = list(g.itertuples(index=False))
        # reward ranking: relevance desc, timestamp desc, then input order
        rank_order = sorted(range(len(rows)),
                            key=lambda j: (-rows[j].relevance, -pd.Timestamp(rows[j].timestamp).value
                                           if not isinstance(rows[j].timestamp, (int, float, np.integer, np.floating))
                                           else -rows[j].timestamp, rows[j].row_))
        rewarded = set(rank_order[:top_k])
        for j, row in enumerate(rows):
            obs.append((float(user), float(row.item_idx)))
            act.append(row.action_)
            rew.append(1.0 if j in rewarded else 0.0)
            last = j == len(rows) - 1
            term.append(1.0 if last else 0.0)
            nxt.append((0.0, 0.0) if last else (float(user), float(rows[j + 1].item_idx)))
    f32 = lambda x, shape: np.asarray(x, dtype=np.float32).reshape(shape)
    return {
        "obs": f32(obs, (n, 2)), "act": f32(act, (n, 1)), "rew": f32(rew, (n, 1)),
        "next_obs": f32(nxt, (n, 2)), "term": f32(term, (n, 1)),
    }
