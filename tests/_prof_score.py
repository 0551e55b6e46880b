import sys; sys.path.insert(0, '.')
import numpy as np, torch
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from replay_cql_b200.synthetic import make_log
from replay_cql_b200.mdp import build_mdp_on_device, seen_csr
dev = torch.device('cuda:0')
log = make_log("ml20m", seed=12345, n_rows=2_000_000)
eng = CqlEngine(CqlHyperParams(batch_size=1024, seed=12345, precision="f16x3"))
build_mdp_on_device(eng, log, top_k=10, action_randomization_scale=1e-3)
U, I = 2048, 26744
users = np.sort(np.random.default_rng(1).choice(138493, size=U, replace=False)).astype(np.int32)
items = np.arange(I, dtype=np.int32)
indptr, seen = seen_csr(log[log["user_idx"].isin(users)], 138493)
d_users, d_items = torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev)
d_ptr, d_seen = torch.from_numpy(indptr).to(dev), torch.from_numpy(seen).to(dev)
oi = torch.empty((U, 10), dtype=torch.int32, device=dev); osc = torch.empty((U, 10), dtype=torch.float32, device=dev)
st = torch.cuda.Stream()
def time_score(tag):
    with torch.cuda.stream(st):
        eng.score_topk_device(d_users[:64], d_items, 10, d_ptr, d_seen, out_items=oi[:64], out_scores=osc[:64], stream=st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.score_topk_device(d_users, d_items, 10, d_ptr, d_seen, out_items=oi, out_scores=osc, stream=st.cuda_stream)
        e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1)
    s = osc.cpu().numpy(); it = oi.cpu().numpy()
    print(f"{tag}: {ms:.1f} ms = {U / ms * 1e3:.0f} users/s; top-1 item id median {np.median(it[:, 0]):.0f}, score range {s.min():.3g}..{s.max():.3g}", flush=True)
time_score("fresh init")
for n in (200, 800, 2000):
    with torch.cuda.stream(st):
        eng.update(n, want_metrics=False, stream=st.cuda_stream)
    st.synchronize()
    time_score(f"after +{n} updates")
eng.close()
