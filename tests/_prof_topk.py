import sys; sys.path.insert(0, '.')
import numpy as np, torch
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
eng = CqlEngine(CqlHyperParams(batch_size=64))
U, I = 8192, 26744
dev = torch.device('cuda:0')
sc = torch.randn((U, I), dtype=torch.float32, device=dev)
users = torch.arange(U, dtype=torch.int32, device=dev)
items = torch.arange(I, dtype=torch.int32, device=dev)
rng = np.random.default_rng(0)
cnt = rng.integers(20, 300, U); indptr = np.zeros(U + 1, np.int64); indptr[1:] = np.cumsum(cnt)
seen = np.concatenate([np.sort(rng.choice(I, c, replace=False)) for c in cnt]).astype(np.int32)
d_ptr = torch.from_numpy(indptr).to(dev); d_seen = torch.from_numpy(seen).to(dev)
# heavy-tailed lists as in bench.py (log-normal history lengths, mean 144, ~3 % of the users above 512 items)
cnt_h = np.minimum(20 + (rng.lognormal(0.0, 1.0, U) * 75).astype(np.int64), I // 2); indptr_h = np.zeros(U + 1, np.int64); indptr_h[1:] = np.cumsum(cnt_h)
seen_h = np.concatenate([np.sort(rng.choice(I, c, replace=False)) for c in cnt_h]).astype(np.int32)
d_ptr_h = torch.from_numpy(indptr_h).to(dev); d_seen_h = torch.from_numpy(seen_h).to(dev)
print('heavy lists: mean', cnt_h.mean(), 'share > 512:', (cnt_h > 512).mean(), 'max', cnt_h.max())
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for variant in ("full", "heavy", "noseen", "noitems"):
        kw = dict(users_t=users, items_t=items, seen_indptr_t=d_ptr, seen_items_t=d_seen)
        if variant == "heavy": kw.update(seen_indptr_t=d_ptr_h, seen_items_t=d_seen_h)
        if variant == "noseen": kw.update(seen_indptr_t=None, seen_items_t=None)
        if variant == "noitems": kw.update(items_t=None, seen_indptr_t=None, seen_items_t=None)
        eng.topk_filter_device(sc, 10, stream=st.cuda_stream, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10): eng.topk_filter_device(sc, 10, stream=st.cuda_stream, **kw)
        e1.record(st); st.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(variant, 'ms', ms, 'GB/s', sc.numel() * 4 / ms / 1e6)
    # pure streaming reference: torch max over rows
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sc.max(dim=1); e0.record(st)
    for _ in range(10): sc.max(dim=1)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('torch rowmax ms', ms, 'GB/s', sc.numel() * 4 / ms / 1e6)
eng.close()
