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
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for variant in ("full", "noseen", "noitems"):
        kw = dict(users_t=users, items_t=items, seen_indptr_t=d_ptr, seen_items_t=d_seen)
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
