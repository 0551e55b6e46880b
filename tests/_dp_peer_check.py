"""torchrun --nproc-per-node N tests/_dp_peer_check.py : data-parallel gradient exchange over NVLink peer memory vs NCCL
(needs N GPUs).  Three engines per rank take the same 6 updates from the same state:
  a: NCCL all-reduce between the four phases (GradAllReducer),
  b: fused exchange -- the update kernels publish / read the gradients themselves; plain `engine.update(6)` (library graph),
  c: the stand-alone peer-memory exchange kernel between the phases (CQL_NO_FUSED_DP path), when requested by env.
N = 2: b == a bitwise (a + b is commutative); any N: replicas bit-identical, no time-outs.  Then timings."""
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
fused = red_b.fused
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(6):
        a.update_data_parallel(red_a, stream=st.cuda_stream)
st.synchronize()
dist.barrier()
if fused:
    b.update(6)                       # the library's own CUDA graph; the kernels exchange the gradients
else:
    with torch.cuda.stream(st):
        for _ in range(6):
            b.update_data_parallel(red_b, stream=st.cuda_stream)
    st.synchronize()
sa, sb = a.get_state(), b.get_state()
same_as_nccl = bool(np.array_equal(sa, sb))
t = torch.from_numpy(sb.astype(np.float64)).cuda()
mx, mn = t.clone(), t.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
replicas_identical = bool(torch.equal(mx, mn))
assert b.get_optimizer()[2] == a.get_optimizer()[2] == 6
print(f"rank {rank}: fused={fused}; peer==nccl bitwise: {same_as_nccl} (expected for world=2); replicas identical: {replicas_identical}; "
      f"timeout flag: {b.dp_error()}; max|diff| {np.abs(sa - sb).max():.3e}", flush=True)
# timing: NCCL (whole DP step as one CUDA graph) vs the peer path
stp = DataParallelStepper(a, reducer=red_a)
stp.run(20); stp.stream.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stp.stream); stp.run(300); e1.record(stp.stream); stp.stream.synchronize()
print(f"rank {rank}: nccl: {e0.elapsed_time(e1) / 300 * 1e3:.1f} us/step (graph={'yes' if stp.graph else 'no'})", flush=True)
stp.graph = None
dist.barrier()
if fused:
    b.update(20); dist.barrier()
    t0 = time.perf_counter(); b.update(300); dt = time.perf_counter() - t0
    print(f"rank {rank}: peer fused (engine.update, library graph): {dt / 300 * 1e6:.1f} us/step, timeout flag {b.dp_error()}", flush=True)
else:
    stp = DataParallelStepper(b, reducer=red_b)
    stp.run(20); stp.stream.synchronize(); dist.barrier()
    e0.record(stp.stream); stp.run(300); e1.record(stp.stream); stp.stream.synchronize()
    print(f"rank {rank}: peer exchange kernel: {e0.elapsed_time(e1) / 300 * 1e3:.1f} us/step, timeout flag {b.dp_error()}", flush=True)
    stp.graph = None
dist.barrier()
torch.cuda.synchronize()
a.close(); b.close()
dist.barrier(); dist.destroy_process_group()
