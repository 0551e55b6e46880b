"""torchrun --nproc-per-node N tests/_dp_peer_check.py : peer-memory gradient exchange vs NCCL (needs N GPUs).
N = 2: bitwise equality with the NCCL path (a + b is commutative); any N: replicas bit-identical, no time-outs."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from replay_cql_b200.parallel import GradAllReducer, PeerGradExchange, DataParallelStepper

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 200000
rng = np.random.default_rng(0)
obs = np.stack([rng.integers(0, 6040, N), rng.integers(0, 3706, N)], 1).astype(np.float32)
act, rew, term = rng.integers(1, 6, N).astype(np.float32), rng.integers(0, 2, N).astype(np.float32), (rng.random(N) < 0.01).astype(np.float32)
def make():
    e = CqlEngine(CqlHyperParams(batch_size=1024, seed=5, precision="f16x3"), device=local, rank=rank, world_size=world)
    e.load_transitions(obs, act, rew, term)
    return e
a, b = make(), make()
red_a = GradAllReducer(a)
try:
    red_b = PeerGradExchange(b)
except Exception as ex:
    print(f"rank {rank}: PeerGradExchange unavailable: {type(ex).__name__}: {ex}", flush=True)
    dist.destroy_process_group(); sys.exit(0)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(6):
        a.update_data_parallel(red_a, stream=st.cuda_stream)
        b.update_data_parallel(red_b, stream=st.cuda_stream)
st.synchronize()
sa, sb = a.get_state(), b.get_state()
same_as_nccl = bool(np.array_equal(sa, sb))
t = torch.from_numpy(sb.astype(np.float64)).cuda()
mx, mn = t.clone(), t.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
replicas_identical = bool(torch.equal(mx, mn))
print(f"rank {rank}: peer==nccl bitwise: {same_as_nccl} (expected for world=2); replicas identical: {replicas_identical}; "
      f"timeout flag: {b.dp_error()}; max|diff| {np.abs(sa - sb).max():.3e}", flush=True)
# timing: whole DP step as one CUDA graph, NCCL vs peer
for name, eng, red in (("nccl", a, red_a), ("peer", b, red_b)):
    stp = DataParallelStepper.__new__(DataParallelStepper)
    stp.engine, stp.reducer, stp.device = eng, red, torch.device("cuda", local)
    stp.stream, stp.graph, stp._warmup, stp.launches_per_step, stp.replayed_steps = torch.cuda.Stream(), None, 3, 0, 0
    stp.run(20); stp.stream.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stp.stream); stp.run(300); e1.record(stp.stream); stp.stream.synchronize()
    ms = e0.elapsed_time(e1) / 300
    print(f"rank {rank}: {name}: {ms * 1e3:.1f} us/step (graph={'yes' if stp.graph else 'no'}), timeout flag {eng.dp_error()}", flush=True)
    dist.barrier()
    stp.graph = None
torch.cuda.synchronize()
a.close(); b.close()
dist.barrier(); dist.destroy_process_group()
