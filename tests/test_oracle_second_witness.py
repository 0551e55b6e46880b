"""CPU: a second witness for the UNPINNED update oracle (oracle/cql_oracle.py).

d3rlpy is not in this container, so the oracle cannot be diffed against the reference.
What CAN be checked is that the oracle's hand-rolled pieces equal the PyTorch building blocks d3rlpy 1.x is made
of -- ``nn.Linear`` + ``ReLU`` encoders, ``torch.distributions.Normal`` + the tanh transform of
``SquashedNormalPolicy``, ``torch.optim.Adam`` with default arguments, ``soft_sync`` Polyak averaging and
``torch.logsumexp`` -- on the same numbers.  A restatement error in ``mlp`` / ``policy_sample`` / ``adam_step`` /
the Polyak step would show up here without needing d3rlpy.
"""
import math

import numpy as np
import pytest
import torch
from torch import nn

from oracle import cql_oracle as O


def _module_from(net, in_f, out_f):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) carrying the oracle net's tensors (d3rlpy VectorEncoder + head)."""
    seq = nn.Sequential(nn.Linear(in_f, O.H), nn.ReLU(), nn.Linear(O.H, O.H), nn.ReLU(), nn.Linear(O.H, out_f))
    seq = seq.to(net["W1"].dtype)
    with torch.no_grad():
        for lin, (w, b) in zip((seq[0], seq[2], seq[4]), (("W1", "b1"), ("W2", "b2"), ("W3", "b3"))):
            lin.weight.copy_(net[w])
            lin.bias.copy_(net[b])
    return seq


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_mlp_equals_nn_linear_stack(dtype):
    gen = torch.Generator().manual_seed(3)
    for in_f, out_f in ((2, 2), (3, 1)):
        net = O.init_net(in_f, out_f, gen, dtype)
        x = torch.randn(97, in_f, generator=gen, dtype=torch.float64).to(dtype) * 50
        ref = _module_from(net, in_f, out_f)(x)
        got = O.mlp(net, x)
        tol = 1e-5 if dtype == torch.float32 else 1e-13   # x @ W.T + b vs the fused addmm: last-ulp differences only
        assert float((got - ref.detach()).abs().max()) <= tol * float(ref.detach().abs().max())


def test_linear_init_is_pytorch_default_bound():
    """nn.Linear default init draws W and b from U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (kaiming_uniform(a=sqrt 5))."""
    gen = torch.Generator().manual_seed(0)
    net = O.init_net(3, 1, gen)
    for w, b, fan_in in ((net["W1"], net["b1"], 3), (net["W2"], net["b2"], O.H), (net["W3"], net["b3"], O.H)):
        bound = 1 / math.sqrt(fan_in)
        assert float(w.abs().max()) <= bound and float(b.abs().max()) <= bound
        if w.numel() > 500:
            assert float(w.abs().max()) > 0.98 * bound      # actually fills the interval
            assert abs(float(w.var()) - bound ** 2 / 3) < 0.05 * bound ** 2 / 3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_policy_sample_equals_torch_distributions(dtype):
    """SquashedNormalPolicy [EXT d3rlpy/models/torch/policies.py]: dist = Normal(mu, logstd.exp()); raw = dist.rsample();
    a = tanh(raw); log_prob = dist.log_prob(raw) - log(1 - a^2 + 1e-6), summed over the action dimension."""
    g = torch.Generator().manual_seed(5)
    B, n = 64, 10
    mu = torch.randn(B, 1, generator=g, dtype=torch.float64).to(dtype)
    logstd = (torch.rand(B, 1, generator=g, dtype=torch.float64) * 6 - 5).to(dtype).clamp(-20, 2)
    eps = torch.randn(B, n, generator=g, dtype=torch.float64).to(dtype)
    a, logp = O.policy_sample(mu, logstd, eps, "eps")
    dist = torch.distributions.Normal(mu, logstd.exp())
    raw = dist.loc + dist.scale * eps            # what rsample() computes from its standard-normal draw
    a_ref = torch.tanh(raw)
    logp_ref = dist.log_prob(raw) - torch.log(1 - a_ref.pow(2) + 1e-6)
    assert torch.equal(a, a_ref)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert float((logp - logp_ref).abs().max()) <= tol * max(1.0, float(logp_ref.abs().max()))
    # the "softplus" variant is the exact log-det of tanh: compare with TanhTransform
    a2, logp2 = O.policy_sample(mu, logstd, eps, "softplus")
    tt = torch.distributions.transforms.TanhTransform()
    logp2_ref = dist.log_prob(raw) - tt.log_abs_det_jacobian(raw, a2)
    assert float((logp2 - logp2_ref).abs().max()) <= tol * max(1.0, float(logp2_ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_adam_step_equals_torch_optim_adam(dtype):
    """Five steps of the oracle's adam_step against torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) on the same
    gradients -- including zero and tiny gradients (the first step moves by lr * g / (|g| + eps))."""
    cfg = O.OracleConfig()
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(300, generator=g, dtype=torch.float64).to(dtype)
    p_or = p0.clone()
    st = {"m": torch.zeros_like(p0), "v": torch.zeros_like(p0)}
    p_t = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_t], lr=cfg.critic_lr, betas=(cfg.beta1, cfg.beta2), eps=cfg.adam_eps)
    for step in range(1, 6):
        grad = torch.randn(300, generator=g, dtype=torch.float64).to(dtype)
        grad[:10] = 0.0
        grad[10:20] *= 1e-10
        O.adam_step(p_or, grad, st, cfg.critic_lr, step, cfg)
        p_t.grad = grad.clone()
        opt.step()
        tol = 2e-7 if dtype == torch.float32 else 1e-15
        assert float((p_or - p_t.detach()).abs().max()) <= tol, step
        s = opt.state[p_t]
        assert float((st["m"] - s["exp_avg"]).abs().max()) <= tol
        assert float((st["v"] - s["exp_avg_sq"]).abs().max()) <= tol


def test_one_update_equals_module_based_restatement():
    """One whole update written a second time with nn.Module networks, torch.optim.Adam (four optimisers), torch
    distributions and autograd ``backward()`` -- the way d3rlpy composes it -- must land on the oracle's losses and
    updated tensors.  Shares no helper with the oracle except the noise and the initial numbers."""
    torch.manual_seed(0)
    cfg = O.OracleConfig()
    st = O.init_state(cfg, seed=13)
    B, n = 48, cfg.n_action_samples
    gen = torch.Generator().manual_seed(1)
    obs = torch.stack([torch.randint(0, 60, (B,), generator=gen).float() * 0.01,
                       torch.randint(0, 37, (B,), generator=gen).float() * 0.01], 1)
    nobs = obs.roll(1, 0)
    term = (torch.rand(B, 1, generator=gen) < 0.1).float()
    batch = {"obs": obs, "act": torch.rand(B, 1, generator=gen), "rew": torch.randint(0, 2, (B, 1), generator=gen).float(),
             "next_obs": nobs * (1 - term), "term": term}
    noise = O.make_noise(B, n, seed=2)

    actor = _module_from(st["actor"], 2, 2)
    critics = [_module_from(c, 3, 1) for c in st["critics"]]
    targ_critics = [_module_from(c, 3, 1) for c in st["targ_critics"]]
    targ_actor = _module_from(st["targ_actor"], 2, 2)
    log_temp = nn.Parameter(st["log_temp"].clone().view(1, 1))
    log_alpha = nn.Parameter(st["log_alpha"].clone().view(1, 1))
    opt_actor = torch.optim.Adam(actor.parameters(), lr=cfg.actor_lr)
    opt_critic = torch.optim.Adam([p for c in critics for p in c.parameters()], lr=cfg.critic_lr)
    opt_temp = torch.optim.Adam([log_temp], lr=cfg.temp_lr)
    opt_alpha = torch.optim.Adam([log_alpha], lr=cfg.alpha_lr)

    def pi(x, eps):                                # SquashedNormalPolicy.sample_with_log_prob / sample_n
        out = actor(x)
        mu, logstd = out[:, 0:1], out[:, 1:2].clamp(-20.0, 2.0)
        dist = torch.distributions.Normal(mu, logstd.exp())
        raw = dist.loc + dist.scale * eps
        a = torch.tanh(raw)
        return a, dist.log_prob(raw) - torch.log(1 - a.pow(2) + 1e-6)

    def q_all(nets, x, a):                         # EnsembleQFunction, reduction "none": [C, rows, 1]
        return torch.stack([net(torch.cat([x, a], 1)) for net in nets], 0)

    def conservative(eps_t, eps_t1, u):
        with torch.no_grad():
            a_t, lp_t = pi(obs, eps_t)
            a_t1, lp_t1 = pi(batch["next_obs"], eps_t1)
        rep = obs.unsqueeze(1).expand(B, n, 2).reshape(B * n, 2)
        v_t = q_all(critics, rep, a_t.reshape(-1, 1)).view(len(critics), B, n) - lp_t.view(1, B, n)
        v_t1 = q_all(critics, rep, a_t1.reshape(-1, 1)).view(len(critics), B, n) - lp_t1.view(1, B, n)
        v_r = q_all(critics, rep, u.reshape(-1, 1)).view(len(critics), B, n) - math.log(0.5 ** cfg.act_dim)
        lse = torch.logsumexp(torch.cat([v_t, v_t1, v_r], 2), dim=2, keepdim=True)
        loss = lse.mean(0).mean() - q_all(critics, obs, batch["act"]).mean(0).mean()
        return (log_alpha.exp().clamp(0, 1e6) * (cfg.conservative_weight * loss - cfg.alpha_threshold)).sum()

    m = {}
    # temp
    opt_temp.zero_grad()
    with torch.no_grad():
        _, lp = pi(obs, noise["temp_eps"])
        targ = lp - cfg.act_dim
    loss = -(log_temp.exp() * targ).mean()
    loss.backward(); opt_temp.step()
    m["temp_loss"], m["temp"] = float(loss.detach()), float(log_temp.detach().exp())
    # alpha
    opt_alpha.zero_grad()
    loss = -conservative(noise["alpha_eps_t"], noise["alpha_eps_t1"], noise["alpha_u"])
    loss.backward(); opt_alpha.step()
    m["alpha_loss"], m["alpha"] = float(loss.detach()), float(log_alpha.detach().exp())
    # critic
    opt_critic.zero_grad()
    with torch.no_grad():
        a1 = torch.tanh(actor(batch["next_obs"])[:, 0:1])
        y = batch["rew"] + cfg.gamma * q_all(targ_critics, batch["next_obs"], a1).min(0).values * (1 - term)
    td = sum(((c(torch.cat([obs, batch["act"]], 1)) - y) ** 2).mean() for c in critics)
    loss = td + conservative(noise["critic_eps_t"], noise["critic_eps_t1"], noise["critic_u"])
    loss.backward(); opt_critic.step()
    m["critic_loss"] = float(loss.detach())
    # actor
    opt_actor.zero_grad()
    a_pi, lp = pi(obs, noise["actor_eps"])
    loss = (log_temp.exp().detach() * lp - q_all(critics, obs, a_pi).min(0).values).mean()
    loss.backward(); opt_actor.step()
    m["actor_loss"] = float(loss.detach())
    # soft_sync
    with torch.no_grad():
        for t, c in zip(targ_critics, critics):
            for pt, pc in zip(t.parameters(), c.parameters()):
                pt.mul_(1 - cfg.tau).add_(cfg.tau * pc)
        for pt, pc in zip(targ_actor.parameters(), actor.parameters()):
            pt.mul_(1 - cfg.tau).add_(cfg.tau * pc)

    m_or, _ = O.update(cfg, st, batch, noise)
    for k in ("temp_loss", "temp", "alpha_loss", "alpha", "critic_loss", "actor_loss"):
        assert abs(m[k] - m_or[k]) <= 1e-5 * max(1.0, abs(m_or[k])), (k, m[k], m_or[k])

    def same(mod, net, tag):
        for lin, (w, b) in zip((mod[0], mod[2], mod[4]), (("W1", "b1"), ("W2", "b2"), ("W3", "b3"))):
            for got, ref in ((lin.weight, net[w]), (lin.bias, net[b])):
                err = float((got.detach() - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
                assert err <= 1e-4, (tag, w, err)     # first Adam step on g ~ 0 entries: lr-sized sign flips allowed
                assert float((got.detach() - ref).abs().median()) <= 1e-7, (tag, w)
    same(actor, st["actor"], "actor")
    same(targ_actor, st["targ_actor"], "targ_actor")
    for i in range(cfg.n_critics):
        same(critics[i], st["critics"][i], f"critic{i}")
        same(targ_critics[i], st["targ_critics"][i], f"targ_critic{i}")
    assert abs(float(log_temp.detach()) - float(st["log_temp"])) <= 1e-7
    assert abs(float(log_alpha.detach()) - float(st["log_alpha"])) <= 1e-7
