"""CPU, world_size 2 over gloo: the N>1 host logic (sharding, gradient averaging, result gather)
and the data-parallel equivalence the trainer relies on: two ranks with half a batch each and
mean-all-reduced gradient groups take the same step as one rank with the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from replay_cql_b200.parallel import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 138_493):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cql_oracle as O
        from replay_cql_b200.parallel import allreduce_mean_, dist_info, gather_rows
        from tests import helpers as Hp
        assert dist_info()[:2] == (rank, world)
        # 0. without NCCL the gradient exchange falls back (collectively) to the all-reduce reducer
        from replay_cql_b200.parallel import GradAllReducer, make_grad_exchange
        assert isinstance(make_grad_exchange(object()), GradAllReducer)
        # 1. allreduce_mean_
        t = torch.full((5,), float(rank + 1))
        allreduce_mean_(t)
        assert torch.allclose(t, torch.full((5,), (1 + world) / 2))
        # 2. ragged gather in rank order
        local = np.full((rank + 2, 3), rank, dtype=np.float64)
        allr = gather_rows(local)
        assert allr.shape == (sum(r + 2 for r in range(world)), 3)
        assert np.all(allr[:2] == 0) and np.all(allr[2:5] == 1)
        # 3. data-parallel step == big-batch step
        B, n = 64, 10
        cfg = O.OracleConfig()
        st = O.init_state(cfg, seed=7, dtype=torch.float64)
        batch = {k: v.double() for k, v in Hp.make_batch(B, seed=1, scale=1e-3).items()}
        noise = {k: v.double() for k, v in O.make_noise(B, n, seed=2).items()}
        lo, hi = shard_range(B, rank, world)
        def hook(gs):
            for g in gs:
                allreduce_mean_(g)
            return gs
        m, _ = O.update(cfg, st, {k: v[lo:hi] for k, v in batch.items()}, {k: v[lo:hi] for k, v in noise.items()},
                        grad_hook=hook)
        flat = torch.cat([st["actor"][k].reshape(-1) for k in O.NET_KEYS] +
                         [c[k].reshape(-1) for c in st["critics"] for k in O.NET_KEYS] +
                         [st["log_temp"].reshape(1), st["log_alpha"].reshape(1)])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered)      # ranks stay bit-identical
        if rank == 0:
            np.save(os.path.join(out_dir, "dp.npy"), flat.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import cql_oracle as O
    from tests import helpers as Hp
    cfg = O.OracleConfig()
    st = O.init_state(cfg, seed=7, dtype=torch.float64)
    batch = {k: v.double() for k, v in Hp.make_batch(64, seed=1, scale=1e-3).items()}
    noise = {k: v.double() for k, v in O.make_noise(64, 10, seed=2).items()}
    O.update(cfg, st, batch, noise)
    ref = torch.cat([st["actor"][k].reshape(-1) for k in O.NET_KEYS] +
                    [c[k].reshape(-1) for c in st["critics"] for k in O.NET_KEYS] +
                    [st["log_temp"].reshape(1), st["log_alpha"].reshape(1)]).numpy()
    dp = np.load(tmp_path / "dp.npy")
    # Adam's first step is lr*sign(g): compare where the full-batch gradient is not ~0
    np.testing.assert_allclose(dp, ref, rtol=0, atol=1e-9)


@pytest.mark.parametrize("world", [2, 3, 4, 8, 16])
def test_fused_exchange_slice_invariants(world):
    """What csrc/dp_peer.cuh relies on (mirrored by parallel.exchange_slices): every value group of a gradient group has
    exactly one owner in [0, world); owners are contiguous and ascending; the two value groups (8 floats) an Adam thread
    updates share an owner; the critic and the actor group (gradient layout [actor | critics | scalars]) both fit."""
    from replay_cql_b200 import layout
    from replay_cql_b200.parallel import exchange_slices
    net = layout.NET_STRIDE
    for off, n in ((0, net), (net, 2 * net), (net, 4 * net)):          # actor; 2 critics; 4 critics
        q_lo, n_q, per, owner = exchange_slices(world, off, n)
        assert per % 2 == 0 and per * world >= n_q
        owners = np.array([owner(q) for q in range(q_lo, q_lo + n_q)])
        assert owners.min() == 0 and owners.max() <= world - 1
        assert np.all(np.diff(owners) >= 0) and np.all(np.diff(owners) <= 1)
        assert np.array_equal(owners[0::2], owners[1::2])               # an 8-float chunk never straddles two owners
        assert (n // 4) % 2 == 0 and (off // 4) % 2 == 0                   # chunks start on even value groups
