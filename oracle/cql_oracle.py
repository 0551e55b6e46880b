"""CPU restatement of one CQL (SAC-based) update.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: follows the published d3rlpy 1.x algorithm as restated in
SURVEY.md Appendix A ([EXT-UNVERIFIED]); upstream locations, for a maintainer
who has d3rlpy at hand:

* ``d3rlpy/algos/cql.py::CQL._update``            -> :func:`update`
* ``algos/torch/sac_impl.py::update_temp``        -> step 1 of :func:`update`
* ``algos/torch/cql_impl.py::update_alpha``       -> step 2
* ``algos/torch/cql_impl.py::compute_critic_loss``
  / ``_compute_conservative_loss``                -> step 3, :func:`conservative`
* ``algos/torch/sac_impl.py::compute_actor_loss`` -> step 4
* ``algos/torch/ddpg_impl.py::update_*_target``   -> step 5 (``soft_sync``)
* ``models/torch/policies.py::SquashedNormalPolicy`` -> :func:`policy_sample`
* ``models/torch/q_functions/mean_q_function.py`` -> :func:`critic_forward`

Everything is plain eager PyTorch on CPU with autograd doing the backward, so
it shares no code and no derivation with the hand-written CUDA backward.  All
random draws are *inputs* (``noise``) so the CUDA path can be fed the same
numbers.  Works in float32 (the parity target) and float64 (error budgeting).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional

import torch

H = 256  # d3rlpy "default" vector encoder: hidden units [256, 256], ReLU


@dataclass
class OracleConfig:
    """Hyper-parameters (SURVEY.md Appendix A defaults)."""

    n_critics: int = 2
    n_action_samples: int = 10
    gamma: float = 0.99
    tau: float = 0.005
    actor_lr: float = 1e-4
    critic_lr: float = 3e-4
    temp_lr: float = 1e-4
    alpha_lr: float = 1e-4
    initial_temperature: float = 1.0
    initial_alpha: float = 1.0
    alpha_threshold: float = 10.0
    conservative_weight: float = 5.0
    soft_q_backup: bool = False  # only the default (False) is restated
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    obs_dim: int = 2
    act_dim: int = 1
    # log|d tanh| term of the squashed Gaussian: "eps" = log(1 - a^2 + 1e-6)
    # (Appendix A); "softplus" = 2(log 2 - x - softplus(-2x)) (later d3rlpy).
    squash: str = "eps"


def _linear_init(out_f: int, in_f: int, gen: torch.Generator, dtype):
    """PyTorch default ``nn.Linear`` init: U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w.to(dtype), b.to(dtype)


def init_net(in_f: int, out_f: int, gen: torch.Generator, dtype=torch.float32):
    """One 3-layer MLP ``in_f -> 256 -> 256 -> out_f`` as a dict of tensors."""
    w1, b1 = _linear_init(H, in_f, gen, dtype)
    w2, b2 = _linear_init(H, H, gen, dtype)
    w3, b3 = _linear_init(out_f, H, gen, dtype)
    return {"W1": w1, "b1": b1, "W2": w2, "b2": b2, "W3": w3, "b3": b3}


NET_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")


def init_state(cfg: OracleConfig, seed: int = 7, dtype=torch.float32) -> Dict:
    """Fresh learner state: actor, critics, hard-synced targets, scalars, Adam."""
    gen = torch.Generator().manual_seed(seed)
    actor = init_net(cfg.obs_dim, 2 * cfg.act_dim, gen, dtype)  # rows: mu, logstd
    critics = [init_net(cfg.obs_dim + cfg.act_dim, 1, gen, dtype) for _ in range(cfg.n_critics)]
    st = {
        "actor": actor,
        "critics": critics,
        "targ_actor": {k: v.clone() for k, v in actor.items()},
        "targ_critics": [{k: v.clone() for k, v in c.items()} for c in critics],
        "log_temp": torch.tensor(math.log(cfg.initial_temperature), dtype=dtype),
        "log_alpha": torch.tensor(math.log(cfg.initial_alpha), dtype=dtype),
        "step": 0,
    }
    st["adam"] = {
        "actor": _adam_zeros(actor),
        "critics": [_adam_zeros(c) for c in critics],
        "log_temp": {"m": torch.zeros((), dtype=dtype), "v": torch.zeros((), dtype=dtype)},
        "log_alpha": {"m": torch.zeros((), dtype=dtype), "v": torch.zeros((), dtype=dtype)},
    }
    return st


def _adam_zeros(net):
    return {k: {"m": torch.zeros_like(v), "v": torch.zeros_like(v)} for k, v in net.items()}


def cast_state(st: Dict, dtype) -> Dict:
    def c(x):
        if isinstance(x, torch.Tensor):
            return x.detach().clone().to(dtype)
        if isinstance(x, dict):
            return {k: c(v) for k, v in x.items()}
        if isinstance(x, list):
            return [c(v) for v in x]
        return x
    return c(st)


# ----------------------------------------------------------------------------- networks
def mlp(net, x):
    h = torch.relu(x @ net["W1"].T + net["b1"])
    h = torch.relu(h @ net["W2"].T + net["b2"])
    return h @ net["W3"].T + net["b3"]


def actor_forward(net, obs):
    """-> (mu, clamped logstd), each [rows, 1]."""
    out = mlp(net, obs)
    return out[:, 0:1], out[:, 1:2].clamp(-20.0, 2.0)


def critic_forward(net, obs, act):
    """Mean Q-function: [rows,2],[rows,1] -> [rows,1]."""
    return mlp(net, torch.cat([obs, act], dim=1))


def policy_sample(mu, logstd, eps, squash: str = "eps"):
    """Squashed-Gaussian rsample with log-prob.  mu/logstd [B,1]; eps [B,n].

    -> actions [B,n], log_probs [B,n] (action dim is 1 so the sum over action
    dims is the identity).
    """
    std = logstd.exp()
    raw = mu + std * eps
    a = torch.tanh(raw)
    normal_logp = -0.5 * ((raw - mu) / std) ** 2 - logstd - 0.5 * math.log(2 * math.pi)
    if squash == "eps":
        logdet = torch.log(1 - a * a + 1e-6)
    elif squash == "softplus":
        logdet = 2 * (math.log(2.0) - raw - torch.nn.functional.softplus(-2 * raw))
    else:
        raise ValueError(squash)
    return a, normal_logp - logdet


def q_ensemble(critics, obs, act):
    return torch.stack([critic_forward(c, obs, act)[:, 0] for c in critics], dim=0)  # [C, rows]


def conservative(cfg, critics, actor, log_alpha, s, a, s1, eps_t, eps_t1, u_rand):
    """``_compute_conservative_loss``; noise [B,n] each, ``u_rand`` in (-1,1)."""
    B, n = eps_t.shape
    with torch.no_grad():
        mu_t, ls_t = actor_forward(actor, s)
        mu_t1, ls_t1 = actor_forward(actor, s1)
        A_t, lp_t = policy_sample(mu_t, ls_t, eps_t, cfg.squash)
        A_t1, lp_t1 = policy_sample(mu_t1, ls_t1, eps_t1, cfg.squash)
    s_rep = s.unsqueeze(1).expand(B, n, s.shape[1]).reshape(B * n, -1)  # value obs is ALWAYS s
    C = len(critics)
    v_t = q_ensemble(critics, s_rep, A_t.reshape(-1, 1)).view(C, B, n) - lp_t.view(1, B, n)
    v_t1 = q_ensemble(critics, s_rep, A_t1.reshape(-1, 1)).view(C, B, n) - lp_t1.view(1, B, n)
    v_r = q_ensemble(critics, s_rep, u_rand.reshape(-1, 1)).view(C, B, n) - math.log(0.5 ** cfg.act_dim)
    lse = torch.logsumexp(torch.cat([v_t, v_t1, v_r], dim=2), dim=2, keepdim=True)  # [C,B,1]
    data = q_ensemble(critics, s, a)  # [C,B]
    loss = lse.mean(dim=0).mean() - data.mean(dim=0).mean()
    scaled = cfg.conservative_weight * loss
    clipped_alpha = log_alpha.exp().clamp(0, 1e6)
    return clipped_alpha * (scaled - cfg.alpha_threshold)


# ----------------------------------------------------------------------------- optimiser
def adam_step(p, g, st, lr, step, cfg):
    """torch.optim.Adam (defaults, no amsgrad / weight decay), one tensor, in place."""
    st["m"].lerp_(g, 1 - cfg.beta1)
    st["v"].mul_(cfg.beta2).addcmul_(g, g, value=1 - cfg.beta2)
    bc1 = 1 - cfg.beta1 ** step
    bc2 = 1 - cfg.beta2 ** step
    denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(cfg.adam_eps)
    p.addcdiv_(st["m"], denom, value=-(lr / bc1))


def _leafify(net):
    return {k: v.detach().clone().requires_grad_(True) for k, v in net.items()}


# ----------------------------------------------------------------------------- the update
NOISE_KEYS = ("temp_eps", "alpha_eps_t", "alpha_eps_t1", "alpha_u",
              "critic_eps_t", "critic_eps_t1", "critic_u", "actor_eps")


def make_noise(B: int, n: int, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """The seven draws of one update (Appendix A 'RNG parity'), as inputs."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).to(dtype)
    ru = lambda *s: (torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1).to(dtype)
    return {
        "temp_eps": rn(B, 1),
        "alpha_eps_t": rn(B, n), "alpha_eps_t1": rn(B, n), "alpha_u": ru(B, n),
        "critic_eps_t": rn(B, n), "critic_eps_t1": rn(B, n), "critic_u": ru(B, n),
        "actor_eps": rn(B, 1),
    }


def update(cfg: OracleConfig, st: Dict, batch: Dict[str, torch.Tensor],
           noise: Dict[str, torch.Tensor], want_grads: bool = False, grad_hook=None):
    """One ``CQL._update(batch)``: temp -> alpha -> critic -> actor -> Polyak.

    ``batch``: obs [B,2], act [B,1], rew [B,1], next_obs [B,2], term [B,1].
    Mutates ``st`` in place.  Returns (metrics dict, grads dict or None).
    ``grad_hook(list_of_grad_tensors) -> list`` (optional) is applied to each of the four
    gradient groups before its Adam step -- the data-parallel tests average over ranks there.
    ``metrics`` keys follow d3rlpy: temp_loss, temp, alpha_loss, alpha,
    critic_loss, actor_loss (temp/alpha are the post-update values).
    """
    assert not cfg.soft_q_backup
    hook = grad_hook if grad_hook is not None else (lambda gs: gs)
    s, a, r, s1, done = (batch[k] for k in ("obs", "act", "rew", "next_obs", "term"))
    st["step"] += 1
    step = st["step"]
    metrics, grads = {}, {}
    actor, critics = st["actor"], st["critics"]

    # 1. temperature ----------------------------------------------------------
    if cfg.temp_lr > 0:
        with torch.no_grad():
            mu, ls = actor_forward(actor, s)
            _, logp = policy_sample(mu, ls, noise["temp_eps"], cfg.squash)
            targ_temp = logp - cfg.act_dim
        lt = st["log_temp"].detach().clone().requires_grad_(True)
        loss = -(lt.exp() * targ_temp).mean()
        (g,) = hook(list(torch.autograd.grad(loss, lt)))
        grads["log_temp"] = g.clone()
        adam_step(st["log_temp"], g, st["adam"]["log_temp"], cfg.temp_lr, step, cfg)
        metrics["temp_loss"] = float(loss.detach())
        metrics["temp"] = float(st["log_temp"].exp())

    # 2. alpha (Lagrange multiplier of the conservative term) ------------------
    if cfg.alpha_lr > 0:
        la = st["log_alpha"].detach().clone().requires_grad_(True)
        loss = -conservative(cfg, critics, actor, la, s, a, s1,
                             noise["alpha_eps_t"], noise["alpha_eps_t1"], noise["alpha_u"])
        (g,) = hook(list(torch.autograd.grad(loss, la)))
        grads["log_alpha"] = g.clone()
        adam_step(st["log_alpha"], g, st["adam"]["log_alpha"], cfg.alpha_lr, step, cfg)
        metrics["alpha_loss"] = float(loss.detach())
        metrics["alpha"] = float(st["log_alpha"].exp())

    # 3. critics ----------------------------------------------------------------
    leaf = [_leafify(c) for c in critics]
    with torch.no_grad():
        mu1, _ = actor_forward(actor, s1)
        a1 = torch.tanh(mu1)  # best_action
        q_targ = q_ensemble(st["targ_critics"], s1, a1).min(dim=0).values.unsqueeze(1)  # [B,1]
        y = r + cfg.gamma * q_targ * (1 - done)
    td = 0.0
    for c in leaf:
        td = td + ((critic_forward(c, s, a) - y) ** 2).mean()
    cons = conservative(cfg, leaf, actor, st["log_alpha"], s, a, s1,
                        noise["critic_eps_t"], noise["critic_eps_t1"], noise["critic_u"])
    loss = td + cons
    flat = [c[k] for c in leaf for k in NET_KEYS]
    gs = hook(list(torch.autograd.grad(loss, flat)))
    grads["critics"] = []
    it = iter(gs)
    for ci, c in enumerate(critics):
        gd = {}
        for k in NET_KEYS:
            g = next(it)
            gd[k] = g.clone()
            adam_step(c[k], g, st["adam"]["critics"][ci][k], cfg.critic_lr, step, cfg)
        grads["critics"].append(gd)
    metrics["critic_loss"] = float(loss.detach())
    metrics["td_loss"] = float(td.detach())

    # 4. actor (through the *updated* critics' inputs) --------------------------
    aleaf = _leafify(actor)
    mu, ls = actor_forward(aleaf, s)
    a_pi, logp = policy_sample(mu, ls, noise["actor_eps"], cfg.squash)
    entropy = st["log_temp"].exp() * logp
    q_min = q_ensemble(critics, s, a_pi).min(dim=0).values.unsqueeze(1)
    loss = (entropy - q_min).mean()
    gs = hook(list(torch.autograd.grad(loss, [aleaf[k] for k in NET_KEYS])))
    grads["actor"] = {}
    for k, g in zip(NET_KEYS, gs):
        grads["actor"][k] = g.clone()
        adam_step(actor[k], g, st["adam"]["actor"][k], cfg.actor_lr, step, cfg)
    metrics["actor_loss"] = float(loss.detach())

    # 5. Polyak ------------------------------------------------------------------
    with torch.no_grad():
        for tc, c in zip(st["targ_critics"], critics):
            for k in NET_KEYS:
                tc[k].mul_(1 - cfg.tau).add_(cfg.tau * c[k])
        for k in NET_KEYS:
            st["targ_actor"][k].mul_(1 - cfg.tau).add_(cfg.tau * actor[k])
    return metrics, (grads if want_grads else None)


# ----------------------------------------------------------------------------- inference
def predict_action(st, obs):
    """d3rlpy ``predict(x)``: greedy action tanh(mu(x))."""
    with torch.no_grad():
        mu, _ = actor_forward(st["actor"], obs)
        return torch.tanh(mu)


def predict_value(st, obs, act):
    """d3rlpy ``predict_value(x, a)``: mean over critics."""
    with torch.no_grad():
        return q_ensemble(st["critics"], obs, act).mean(dim=0)


def relevance(st, obs, mode: str = "q"):
    """RePlay-wrapper relevance (Appendix B): 'q' = predict_value(x, predict(x)); 'policy' = predict(x)."""
    a = predict_action(st, obs)
    if mode == "policy":
        return a[:, 0]
    return predict_value(st, obs, a)
