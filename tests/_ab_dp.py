"""Scratch: torchrun --nproc-per-node N tests/_ab_dp.py [steps] -- fused data-parallel update timing (library graph)."""
import os, sys, time; sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from replay_cql_b200.engine import CqlEngine, CqlHyperParams
from replay_cql_b200.parallel import make_grad_exchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = CqlEngine(CqlHyperParams(batch_size=1024, seed=12345, precision="f16x3"), device=local, rank=rank, world_size=world)
eng.synth_table(20_000_263, 138_493, 26_744, seed=12345)
red = None if os.environ.get("AB_NODP") else make_grad_exchange(eng)
st = torch.cuda.Stream()
dist.barrier()
with torch.cuda.stream(st):
    eng.update(50, want_metrics=False, stream=st.cuda_stream)
    torch.cuda.synchronize(); dist.barrier()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier()
        e0.record(st); eng.update(steps, want_metrics=False, stream=st.cuda_stream); e1.record(st)
        torch.cuda.synchronize()
        mine = e0.elapsed_time(e1) / steps
        if os.environ.get("AB_NODP"): print(f"rank {rank} rep {rep}: {mine * 1e3:.2f} us/step (independent replicas, no exchange)", flush=True)
        t = torch.tensor([mine], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
if rank == 0:
    print(os.environ.get("AB_TAG", "-"), f"world {world} fused={getattr(red, 'fused', None)} us/step {best * 1e3:.2f} updates/s {world * 1e3 / best:.0f} err={eng.dp_error()}", flush=True)
torch.cuda.synchronize(); dist.barrier()
eng.close()
dist.barrier(); dist.destroy_process_group()
