"""Scratch: which path (graph / phase split / single-step) diverges under the early PDL trigger.  Usage:
python tests/_diag_trigger.py <tag>   -> gpurun_out/trig_<tag>.npz   (states after 5 updates, per precision and path)"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from replay_cql_b200.synthetic import make_log
from replay_cql_b200.mdp import build_mdp
tag = sys.argv[1]
log = make_log("tiny"); mdp = build_mdp(log, seed=1)
out = {}
for prec in ("fp32", "f16x3"):
    for path in ("graph", "phase", "single", "phase_sync"):
        e = CqlEngine(CqlHyperParams(batch_size=64, seed=9, precision=prec), device=0)
        e.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
        if path == "graph": e.update(5)
        elif path == "single":
            for _ in range(5): e.update(1); torch.cuda.synchronize()
        elif path == "phase":
            for _ in range(5): e.update_data_parallel(lambda buf: None)
        else:
            for _ in range(5): e.update_data_parallel(lambda buf: torch.cuda.synchronize())
        torch.cuda.synchronize()
        out[f"{prec}_{path}"] = e.get_state().copy()
        e.close()
np.savez(f"gpurun_out/trig_{tag}.npz", **out)
for prec in ("fp32", "f16x3"):
    base = out[f"{prec}_graph"]
    print(tag, prec, {p: int((out[f"{prec}_{p}"] != base).sum()) for p in ("phase", "single", "phase_sync")})
