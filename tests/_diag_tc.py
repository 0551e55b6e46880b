import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import cql_oracle as O
from replay_cql_b200 import layout
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from tests import helpers as Hp
B = 256
for scale in (1e-3, 1.0):
    cfg = O.OracleConfig(); st = O.init_state(cfg, seed=7)
    flat = Hp.oracle_state_to_flat(st)
    batch = Hp.make_batch(B, seed=100, scale=scale); noise = O.make_noise(B, 10, seed=200)
    res = {}
    for prec in ("fp32", "tf32x3", "bf16"):
        eng = CqlEngine(CqlHyperParams(batch_size=B, precision=prec))
        eng.set_state(flat)
        m, g = eng.update_batch(Hp.batch_to_numpy(batch), Hp.noise_to_numpy(noise), True)
        res[prec] = (m, g, eng.get_state())
        eng.close()
        print(scale, prec, {k: round(v, 5) for k, v in m.items()}, flush=True)
    for prec in ("tf32x3", "bf16"):
        m, g, s = res[prec]; m0, g0, s0 = res["fp32"]
        print(' ', prec, 'metrics rel', max(abs(m[k]-m0[k])/max(1,abs(m0[k])) for k in m),
              'critic grad W2 %.2e' % Hp.rel_err(g['critics'][0]['W2'], g0['critics'][0]['W2']),
              'actor grad W2 %.2e' % Hp.rel_err(g['actor']['W2'], g0['actor']['W2']),
              'state %.2e' % Hp.rel_err(s, s0), flush=True)
        print('     grads', {k: ('%.1e' % Hp.rel_err(g['actor'][k], g0['actor'][k]), '%.1e' % Hp.rel_err(g['critics'][0][k], g0['critics'][0][k]), '%.1e' % Hp.rel_err(g['critics'][1][k], g0['critics'][1][k])) for k in layout.NET_KEYS}, flush=True)
N = 200000
rng = np.random.default_rng(0)
obs = np.stack([rng.integers(0, 6040, N), rng.integers(0, 3706, N)], 1).astype(np.float32)
for prec in ("fp32", "tf32x3", "bf16"):
    eng = CqlEngine(CqlHyperParams(batch_size=1024, precision=prec))
    eng.load_transitions(obs, rng.integers(1, 6, N).astype(np.float32), rng.integers(0, 2, N).astype(np.float32), (rng.random(N) < 0.01).astype(np.float32))
    eng.update(5)
    t = time.time(); m = eng.update(100); dt = time.time() - t
    print(prec, 'updates/s %.1f' % (100 / dt), {k: round(v, 4) for k, v in m.items()}, flush=True)
    tk = [eng.timed_update() for _ in range(4)][1:]
    print('   ', {k: round(float(np.mean([t[k] for t in tk])), 4) for k in tk[0]}, flush=True)
    eng.close()
