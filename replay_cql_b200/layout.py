"""Flat parameter layout shared with the kernels (``csrc/common.cuh``).

State = ``[actor | critic_0..C-1 | targ_actor | targ_critic_0..C-1 | scalars(64)]``;
each network owns a ``NET_STRIDE``-float slot laid out
``W1[H][in] b1[H] W2[H][H] b2[H] W3[out][H] b3[out]`` in PyTorch's (out, in)
row-major convention, so a d3rlpy/torch ``state_dict`` maps onto it directly.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

H = 256
NET_STRIDE = 67136
SCALAR_SLOT = 64
NET_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")
ACTOR_IN, ACTOR_OUT = 2, 2   # obs -> (mu, logstd)
CRITIC_IN, CRITIC_OUT = 3, 1  # (obs, act) -> q


def net_shapes(in_dim: int, out_dim: int):
    return {"W1": (H, in_dim), "b1": (H,), "W2": (H, H), "b2": (H,), "W3": (out_dim, H), "b3": (out_dim,)}


def net_floats(in_dim: int, out_dim: int) -> int:
    return sum(int(np.prod(s)) for s in net_shapes(in_dim, out_dim).values())


def state_floats(n_critics: int) -> int:
    return (2 + 2 * n_critics) * NET_STRIDE + SCALAR_SLOT


def grad_floats(n_critics: int) -> int:
    return (1 + n_critics) * NET_STRIDE + SCALAR_SLOT


def slot_actor() -> int:
    return 0


def slot_critic(c: int) -> int:
    return 1 + c


def slot_targ_actor(n_critics: int) -> int:
    return 1 + n_critics


def slot_targ_critic(n_critics: int, c: int) -> int:
    return 2 + n_critics + c


def scalars_off(n_critics: int) -> int:
    return (2 + 2 * n_critics) * NET_STRIDE


def pack_net(net: Dict[str, np.ndarray], in_dim: int, out_dim: int, out: np.ndarray) -> None:
    """Write one network dict into its ``NET_STRIDE`` slot ``out`` (zero padded)."""
    out[:] = 0.0
    pos = 0
    for key, shape in net_shapes(in_dim, out_dim).items():
        arr = np.asarray(net[key], dtype=np.float32)
        if arr.shape != shape:
            raise ValueError(f"{key}: expected shape {shape}, got {arr.shape}")
        cnt = arr.size
        out[pos:pos + cnt] = arr.reshape(-1)
        pos += cnt


def unpack_net(slot: np.ndarray, in_dim: int, out_dim: int) -> Dict[str, np.ndarray]:
    net, pos = {}, 0
    for key, shape in net_shapes(in_dim, out_dim).items():
        cnt = int(np.prod(shape))
        net[key] = np.array(slot[pos:pos + cnt], dtype=np.float32).reshape(shape)
        pos += cnt
    return net


def pack_state(actor, critics: List[dict], targ_actor, targ_critics: List[dict],
               log_temp: float, log_alpha: float) -> np.ndarray:
    C = len(critics)
    flat = np.zeros(state_floats(C), dtype=np.float32)
    s = lambda i: flat[i * NET_STRIDE:(i + 1) * NET_STRIDE]
    pack_net(actor, ACTOR_IN, ACTOR_OUT, s(slot_actor()))
    pack_net(targ_actor, ACTOR_IN, ACTOR_OUT, s(slot_targ_actor(C)))
    for c in range(C):
        pack_net(critics[c], CRITIC_IN, CRITIC_OUT, s(slot_critic(c)))
        pack_net(targ_critics[c], CRITIC_IN, CRITIC_OUT, s(slot_targ_critic(C, c)))
    flat[scalars_off(C)] = log_temp
    flat[scalars_off(C) + 1] = log_alpha
    return flat


def unpack_state(flat: np.ndarray, n_critics: int) -> dict:
    C = n_critics
    s = lambda i: flat[i * NET_STRIDE:(i + 1) * NET_STRIDE]
    return {
        "actor": unpack_net(s(slot_actor()), ACTOR_IN, ACTOR_OUT),
        "critics": [unpack_net(s(slot_critic(c)), CRITIC_IN, CRITIC_OUT) for c in range(C)],
        "targ_actor": unpack_net(s(slot_targ_actor(C)), ACTOR_IN, ACTOR_OUT),
        "targ_critics": [unpack_net(s(slot_targ_critic(C, c)), CRITIC_IN, CRITIC_OUT) for c in range(C)],
        "log_temp": float(flat[scalars_off(C)]),
        "log_alpha": float(flat[scalars_off(C) + 1]),
    }


def unpack_grads(flat: np.ndarray, n_critics: int) -> dict:
    """Trainable-gradient buffer ``[actor | critics | scalars]`` -> named arrays."""
    C = n_critics
    s = lambda i: flat[i * NET_STRIDE:(i + 1) * NET_STRIDE]
    so = (1 + C) * NET_STRIDE
    return {
        "actor": unpack_net(s(0), ACTOR_IN, ACTOR_OUT),
        "critics": [unpack_net(s(1 + c), CRITIC_IN, CRITIC_OUT) for c in range(C)],
        "log_temp": float(flat[so]),
        "log_alpha": float(flat[so + 1]),
    }


def _linear_init(rng: np.random.Generator, out_f: int, in_f: int):
    """PyTorch default ``nn.Linear`` init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for W and b."""
    bound = 1.0 / math.sqrt(in_f)
    w = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return w, b


def init_net(rng: np.random.Generator, in_dim: int, out_dim: int) -> Dict[str, np.ndarray]:
    w1, b1 = _linear_init(rng, H, in_dim)
    w2, b2 = _linear_init(rng, H, H)
    w3, b3 = _linear_init(rng, out_dim, H)
    return {"W1": w1, "b1": b1, "W2": w2, "b2": b2, "W3": w3, "b3": b3}


def init_state(n_critics: int, seed: int, initial_temperature: float = 1.0,
               initial_alpha: float = 1.0) -> np.ndarray:
    """Fresh flat state: random actor/critics, hard-synced targets, log scalars."""
    rng = np.random.default_rng(seed)
    actor = init_net(rng, ACTOR_IN, ACTOR_OUT)
    critics = [init_net(rng, CRITIC_IN, CRITIC_OUT) for _ in range(n_critics)]
    return pack_state(actor, critics, actor, critics, math.log(initial_temperature), math.log(initial_alpha))
