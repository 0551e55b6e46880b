import sys; sys.path.insert(0, '.')
import numpy as np
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
N = 200000
rng = np.random.default_rng(0)
obs = np.stack([rng.integers(0, 6040, N), rng.integers(0, 3706, N)], 1).astype(np.float32)
eng = CqlEngine(CqlHyperParams(batch_size=1024, precision=prec))
eng.load_transitions(obs, rng.integers(1, 6, N).astype(np.float32), rng.integers(0, 2, N).astype(np.float32), (rng.random(N) < 0.01).astype(np.float32))
for _ in range(4):
    eng.timed_update()
print(eng.timed_update())
eng.close()
