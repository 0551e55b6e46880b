"""Scratch A/B harness: HBM-resident updates/s (graph replay, batch 1024, 20M-row synthetic table) + per-kernel event times.
python tests/_ab_update.py [steps] ; environment knobs (CQL_*) select the variant; CQL_LIB another build."""
import sys, os, json; sys.path.insert(0, '.')
import numpy as np, torch
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(os.environ.get("AB_BATCH", 1024))
eng = CqlEngine(CqlHyperParams(batch_size=B, seed=12345, precision=os.environ.get("AB_PREC", "f16x3")), device=0)
eng.synth_table(20_000_263, 138_493, 26_744, seed=12345)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    eng.update(50, want_metrics=False, stream=st.cuda_stream)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); eng.update(steps, want_metrics=False, stream=st.cuda_stream); e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
t = [eng.timed_update(stream=st.cuda_stream) for _ in range(5)][1:]
tk = {k: round(float(np.mean([x[k] for x in t])) * 1e3, 1) for k in t[0]}
print(os.environ.get("AB_TAG", "-"), "us/step %.2f  updates/s %.0f " % (best * 1e3, 1e3 / best), tk, eng.read_metrics()["critic_loss"])
eng.close()
