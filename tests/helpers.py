"""Shared test helpers: oracle <-> flat-state conversion, seeded batches."""
from __future__ import annotations

import numpy as np
import torch

from oracle import cql_oracle as O
from replay_cql_b200 import layout

NET_KEYS = layout.NET_KEYS


def _np(net):
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in net.items()}


def oracle_state_to_flat(st) -> np.ndarray:
    return layout.pack_state(_np(st["actor"]), [_np(c) for c in st["critics"]], _np(st["targ_actor"]),
                             [_np(c) for c in st["targ_critics"]], float(st["log_temp"]), float(st["log_alpha"]))


def oracle_adam_to_flat(st):
    """-> (m, v, step) in the flat state layout (targets zero)."""
    C = len(st["critics"])
    m = np.zeros(layout.state_floats(C), dtype=np.float32)
    v = np.zeros_like(m)
    def put(buf, key, slot, net_adam, i, o):
        d = {k: net_adam[k][key].numpy().astype(np.float32) for k in NET_KEYS}
        layout.pack_net(d, i, o, buf[slot * layout.NET_STRIDE:(slot + 1) * layout.NET_STRIDE])
    for buf, key in ((m, "m"), (v, "v")):
        put(buf, key, layout.slot_actor(), st["adam"]["actor"], layout.ACTOR_IN, layout.ACTOR_OUT)
        for c in range(C):
            put(buf, key, layout.slot_critic(c), st["adam"]["critics"][c], layout.CRITIC_IN, layout.CRITIC_OUT)
        so = layout.scalars_off(C)
        buf[so] = float(st["adam"]["log_temp"][key])
        buf[so + 1] = float(st["adam"]["log_alpha"][key])
    return m, v, int(st["step"])


def make_batch(B: int, seed: int, n_users: int = 6040, n_items: int = 3706, scale: float = 1.0,
               term_frac: float = 0.05):
    """Seeded synthetic minibatch.  ``scale`` < 1 shrinks observations so that the policy is NOT
    saturated (raw indices drive tanh to +-1, which hides most of the actor math)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.randint(0, n_users, (B,), generator=g).float() * scale
    i = torch.randint(0, n_items, (B,), generator=g).float() * scale
    i2 = torch.randint(0, n_items, (B,), generator=g).float() * scale
    term = (torch.rand(B, 1, generator=g) < term_frac).float()
    act = torch.randint(1, 6, (B, 1), generator=g).float() + 1e-3 * torch.randn(B, 1, generator=g)
    if scale != 1.0:
        act = act / 5.0
    batch = {
        "obs": torch.stack([u, i], 1), "act": act,
        "rew": torch.randint(0, 2, (B, 1), generator=g).float(),
        "next_obs": torch.stack([u, i2], 1) * (1 - term), "term": term,
    }
    return batch


def batch_to_numpy(batch):
    return {k: v.numpy().astype(np.float32) for k, v in batch.items()}


def noise_to_numpy(noise):
    return {k: v.numpy().astype(np.float32) for k, v in noise.items()}


def rel_err(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def flat_to_oracle_state(flat: np.ndarray, cfg, dtype=torch.float32):
    """Flat product-layout state -> fresh oracle state dict (zero Adam moments, step 0)."""
    un = layout.unpack_state(np.asarray(flat, dtype=np.float32), cfg.n_critics)
    t = lambda net: {k: torch.from_numpy(v.copy()).to(dtype) for k, v in net.items()}
    st = O.init_state(cfg, seed=0, dtype=dtype)
    st["actor"], st["targ_actor"] = t(un["actor"]), t(un["targ_actor"])
    st["critics"] = [t(c) for c in un["critics"]]
    st["targ_critics"] = [t(c) for c in un["targ_critics"]]
    st["log_temp"] = torch.tensor(un["log_temp"], dtype=dtype)
    st["log_alpha"] = torch.tensor(un["log_alpha"], dtype=dtype)
    return st


def digest(flat: np.ndarray, n_probe: int = 512, seed: int = 99) -> np.ndarray:
    """Compact fingerprint of a flat buffer: per-network-slot L2 norm and sum + probed entries."""
    flat = np.asarray(flat, dtype=np.float64)
    n_slots = flat.size // layout.NET_STRIDE
    parts = []
    for s in range(n_slots):
        sl = flat[s * layout.NET_STRIDE:(s + 1) * layout.NET_STRIDE]
        parts += [np.sqrt((sl * sl).sum()), sl.sum()]
    parts += list(flat[n_slots * layout.NET_STRIDE:n_slots * layout.NET_STRIDE + 2])
    idx = np.random.default_rng(seed).integers(0, n_slots * layout.NET_STRIDE, size=n_probe)
    return np.concatenate([np.asarray(parts), flat[idx]])


def rel_err_l2(a, b) -> float:
    """||a - b||_2 / ||b||_2 (Frobenius)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.sqrt((b * b).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return float(num / den) if den > 0 else float(num)
