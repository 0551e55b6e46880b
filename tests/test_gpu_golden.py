"""GPU: the CUDA update against the committed golden vectors (no oracle run needed)."""
from pathlib import Path

import numpy as np
import pytest

from replay_cql_b200 import layout
from tests import helpers as Hp

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
NAMES = ("temp_loss", "temp", "alpha_loss", "alpha", "critic_loss", "actor_loss")


@pytest.mark.parametrize("name", ["update_scaled_eps.npz", "update_scaled_softplus.npz", "update_rawidx_eps.npz"])
def test_update_matches_golden(engine_factory, name):
    z = np.load(GOLD / name)
    B = int(z["B"])
    eng = engine_factory(batch_size=B, squash=str(z["squash"]))
    eng.set_state(layout.init_state(2, int(z["init_seed"])))
    for s in range(int(z["steps"])):
        batch = {k: z[f"batch{s}_{k}"] for k in ("obs", "act", "rew", "next_obs", "term")}
        noise = {k: z[f"noise{s}_{k}"] for k in ("temp_eps", "alpha_eps_t", "alpha_eps_t1", "alpha_u",
                                                 "critic_eps_t", "critic_eps_t1", "critic_u", "actor_eps")}
        m, g = eng.update_batch(batch, noise, want_grads=True)
        got = np.array([m[k] for k in NAMES])
        ref = z[f"metrics{s}"]
        assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-4, (s, got, ref)
        # state digest: slot norms / sums / probed entries of the updated parameters
        sd = Hp.digest(eng.get_state())
        ref_sd = z[f"state_digest{s}"]
        n_slots = 6
        assert np.max(np.abs(sd[:2 * n_slots:2] - ref_sd[:2 * n_slots:2]) / ref_sd[:2 * n_slots:2]) <= 1e-4
        probe = slice(2 * n_slots + 2, None)
        assert np.max(np.abs(sd[probe] - ref_sd[probe])) <= 1e-4 * np.max(np.abs(ref_sd[probe]))
