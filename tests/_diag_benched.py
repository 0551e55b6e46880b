"""scratch diagnostic (not a test): where do f16x3 / fp32 differ from the oracle on raw-index inputs?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import cql_oracle as O
from replay_cql_b200 import layout
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from tests import helpers as Hp

B, nu, ni = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 138_493, 26_744
if B == 8192:
    nu, ni = 10_000_000, 1_000_000
cfg = O.OracleConfig()
for precision in ("f16x3", "fp32"):
    st = O.init_state(cfg, seed=7)
    eng = CqlEngine(CqlHyperParams(batch_size=B, precision=precision), device=0)
    for step in range(3):
        batch = Hp.make_batch(B, seed=500 + step, n_users=nu, n_items=ni, scale=1.0)
        noise = O.make_noise(B, cfg.n_action_samples, seed=600 + step)
        eng.set_state(Hp.oracle_state_to_flat(st))
        eng.set_optimizer(*Hp.oracle_adam_to_flat(st))
        st64 = O.cast_state(st, torch.float64)
        m64, g64 = O.update(cfg, st64, {k: v.double() for k, v in batch.items()}, {k: v.double() for k, v in noise.items()}, want_grads=True)
        m32, g32 = O.update(cfg, st, batch, noise, want_grads=True)
        mg, gg = eng.update_batch(Hp.batch_to_numpy(batch), Hp.noise_to_numpy(noise), want_grads=True)
        print(f"== {precision} step {step}")
        for k in m32:
            if k in mg:
                print(f"  {k:12s} gpu {mg[k]:.8g} o32 {m32[k]:.8g} o64 {m64[k]:.8g}")
        un_g = layout.unpack_state(eng.get_state(), 2)
        un_32 = layout.unpack_state(Hp.oracle_state_to_flat(st), 2)
        for grp, gget in (("actor", lambda g, k: g["actor"][k]), ("critics0", lambda g, k: g["critics"][0][k]), ("critics1", lambda g, k: g["critics"][1][k])):
            for k in layout.NET_KEYS:
                a = np.asarray(gget(gg, k), dtype=np.float64)
                b32 = gget(g32, k).numpy().astype(np.float64)
                b64 = gget(g64, k).numpy()
                mx = np.abs(b64).max()
                wg = un_g["actor"][k] if grp == "actor" else un_g["critics"][int(grp[-1])][k]
                w32 = un_32["actor"][k] if grp == "actor" else un_32["critics"][int(grp[-1])][k]
                werr = np.abs(wg.astype(np.float64) - w32).max() / np.abs(w32).max()
                e = np.abs(a - b64)
                i = np.unravel_index(np.argmax(e), e.shape)
                wi = np.unravel_index(np.argmax(np.abs(wg.astype(np.float64) - w32)), wg.shape)
                print(f"  {grp:8s} {k:3s} max|g| {mx:.3e} grad err gpu-64 {e.max()/mx:.2e} (o32-64 {np.abs(b32-b64).max()/mx:.2e}) nnz {np.count_nonzero(b64)}/{b64.size} | W err {werr:.2e} at {wi}: g gpu {a[wi]:.4e} o32 {b32[wi]:.4e} o64 {b64[wi]:.4e}")
    eng.close()
