"""Generates the committed golden vectors under tests/golden/.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

"self-oracle" vectors (the reference holds no golden vector for
this path, so these pin OUR restatement -- parity stays "unpinned" against d3rlpy):

* update_*.npz : init seed, batch, noise-in -> 6 losses, digests (norms, sums, 512 probed entries)
                 of all gradients and of state-out (float32 oracle) + the float64 twin's losses for error budgeting
* score_*.npz  : weights, users, items, seen CSR -> top-k lists and scores
* mdp_ref_fixture.npz : the reference's own test log (tests/utils.py:59-76) -> transitions
"""
from __future__ import annotations

import sys
from datetime import datetime
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import cql_oracle as O          # noqa: E402
from oracle import mdp_oracle, recs_oracle  # noqa: E402
from tests import helpers as Hp             # noqa: E402
from replay_cql_b200 import layout          # noqa: E402

OUT = ROOT / "tests" / "golden"

REF_LOG = [  # /root/reference/tests/utils.py:59-76
    [0, 0, datetime(2019, 8, 22), 4.0], [0, 2, datetime(2019, 8, 23), 3.0], [0, 1, datetime(2019, 8, 27), 2.0],
    [1, 3, datetime(2019, 8, 24), 3.0], [1, 0, datetime(2019, 8, 25), 4.0], [2, 1, datetime(2019, 8, 26), 5.0],
    [2, 0, datetime(2019, 8, 26), 5.0], [2, 2, datetime(2019, 8, 26), 3.0], [3, 1, datetime(2019, 8, 26), 5.0],
    [3, 0, datetime(2019, 8, 26), 5.0], [3, 0, datetime(2019, 8, 26), 1.0],
]


def flat_grads(g, C):
    out = np.zeros(layout.grad_floats(C), dtype=np.float32)
    layout.pack_net({k: v.numpy() for k, v in g["actor"].items()}, 2, 2, out[:layout.NET_STRIDE])
    for c in range(C):
        layout.pack_net({k: v.numpy() for k, v in g["critics"][c].items()}, 3, 1,
                        out[(1 + c) * layout.NET_STRIDE:(2 + c) * layout.NET_STRIDE])
    so = (1 + C) * layout.NET_STRIDE
    out[so], out[so + 1] = float(g["log_temp"]), float(g["log_alpha"])
    return out


def golden_update(name: str, B: int, scale: float, squash: str, steps: int = 2):
    cfg = O.OracleConfig(squash=squash)
    st = Hp.flat_to_oracle_state(layout.init_state(cfg.n_critics, 7), cfg)   # weights = f(seed): not stored
    st64 = O.cast_state(st, torch.float64)
    rec = {"init_seed": 7, "B": B, "scale": scale, "squash": squash, "steps": steps}
    for s in range(steps):
        batch = Hp.make_batch(B, seed=100 + s, scale=scale)
        noise = O.make_noise(B, cfg.n_action_samples, seed=200 + s)
        m, g = O.update(cfg, st, batch, noise, want_grads=True)
        b64 = {k: v.double() for k, v in batch.items()}
        n64 = {k: v.double() for k, v in noise.items()}
        m64, _ = O.update(cfg, st64, b64, n64)
        for k, v in Hp.batch_to_numpy(batch).items():
            rec[f"batch{s}_{k}"] = v
        for k, v in Hp.noise_to_numpy(noise).items():
            rec[f"noise{s}_{k}"] = v
        names = ("temp_loss", "temp", "alpha_loss", "alpha", "critic_loss", "actor_loss")
        rec[f"metrics{s}"] = np.array([m[k] for k in names], dtype=np.float64)
        rec[f"metrics64_{s}"] = np.array([m64[k] for k in names], dtype=np.float64)
        rec[f"grads_digest{s}"] = Hp.digest(flat_grads(g, cfg.n_critics))
        rec[f"state_digest{s}"] = Hp.digest(Hp.oracle_state_to_flat(st))
    np.savez_compressed(OUT / name, **rec)


def golden_score(name: str, U: int, I: int, k: int, seed: int, scale: float = 1.0):
    cfg = O.OracleConfig()
    st = Hp.flat_to_oracle_state(layout.init_state(cfg.n_critics, seed), cfg)
    rng = np.random.default_rng(seed)
    users = np.sort(rng.choice(500, size=U, replace=False)).astype(np.int32)
    items = np.sort(rng.choice(4000, size=I, replace=False)).astype(np.int32)
    seen = {int(u): set(rng.choice(items, size=rng.integers(0, min(40, I)), replace=False).tolist()) for u in users}
    def score(obs):
        return O.relevance(st, torch.from_numpy(obs), "q").numpy()
    def policy(obs):
        return O.relevance(st, torch.from_numpy(obs), "policy").numpy()
    ti, ts = recs_oracle.brute_force_topk(score, users, items, seen, k)
    pi, ps = recs_oracle.brute_force_topk(policy, users, items, seen, k)
    indptr = np.zeros(int(users.max()) + 2, dtype=np.int64)
    flat = []
    for u in range(int(users.max()) + 1):
        s = sorted(seen.get(u, ()))
        flat.extend(s)
        indptr[u + 1] = len(flat)
    np.savez_compressed(OUT / name, init_seed=seed, users=users, items=items, k=k,
                        seen_indptr=indptr, seen_items=np.asarray(flat, dtype=np.int32),
                        top_items=ti, top_scores=ts, pol_items=pi, pol_scores=ps)


def golden_mdp():
    log = pd.DataFrame(REF_LOG, columns=["user_idx", "item_idx", "timestamp", "relevance"])
    noise = np.random.default_rng(0).standard_normal(len(log)) * 1e-3
    out = mdp_oracle.build_mdp(log, top_k=1, action_noise=noise)
    np.savez_compressed(OUT / "mdp_ref_fixture.npz", action_noise=noise, **out)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    golden_update("update_scaled_eps.npz", B=64, scale=1e-3, squash="eps")
    golden_update("update_scaled_softplus.npz", B=64, scale=1e-3, squash="softplus")
    golden_update("update_rawidx_eps.npz", B=64, scale=1.0, squash="eps", steps=1)
    golden_score("score_small.npz", U=7, I=333, k=10, seed=3)
    golden_score("score_k1.npz", U=3, I=70, k=1, seed=4)
    golden_mdp()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
