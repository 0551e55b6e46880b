import sys; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import cql_oracle as O
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from tests import helpers as Hp
B = 256
for scale, sb, sn in ((1.0, 300, 400), (1.0, 301, 401)):
    cfg = O.OracleConfig(); st = O.init_state(cfg, seed=7)
    flat = Hp.oracle_state_to_flat(st)
    st64 = O.cast_state(st, torch.float64)
    batch = Hp.make_batch(B, seed=sb, scale=scale); noise = O.make_noise(B, 10, seed=sn)
    m64, g64 = O.update(cfg, st64, {k: v.double() for k, v in batch.items()}, {k: v.double() for k, v in noise.items()}, True)
    for prec in ("fp32", "tf32x3"):
        eng = CqlEngine(CqlHyperParams(batch_size=B, precision=prec)); eng.set_state(flat)
        m, g = eng.update_batch(Hp.batch_to_numpy(batch), Hp.noise_to_numpy(noise), True); eng.close()
        for c in range(2):
            ref = g64['critics'][c]['W2'].numpy(); got = g['critics'][c]['W2'].astype(np.float64)
            err = np.abs(got - ref); mx = np.abs(ref).max()
            rowmax = err.max(axis=1) / mx
            top = np.argsort(-rowmax)[:4]
            print(scale, prec, 'critic', c, 'W2 max rel %.2e  median row err %.2e  top rows' % (err.max() / mx, np.median(rowmax)), [(int(j), '%.1e' % rowmax[j]) for j in top],
                  ' b2 err rows', ['%.1e' % (abs(g['critics'][c]['b2'][j] - g64['critics'][c]['b2'].numpy()[j]) / np.abs(g64['critics'][c]['b2'].numpy()).max()) for j in top], flush=True)
