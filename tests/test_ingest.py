"""Spark-free ingestion and the user-range split of the MDP build (8e).

CPU part: chunk enumeration of pyarrow / pandas columns (no intermediate frame), the user-range split (balanced by
rows, episodes never cut), and the split under a world-size-2 gloo group.  GPU part: a pyarrow.Table with several
chunks per column, a Parquet file with several row groups, odd dtypes and timestamp types build the replay table
bit-exactly equal to the host builder's (which the loop oracle pins, tests/test_mdp.py).
"""
import os
import socket

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from replay_cql_b200 import mdp
from replay_cql_b200.synthetic import make_log


def _table(log, n_chunks=3):
    t = pa.Table.from_pandas(log, preserve_index=False)
    cuts = np.linspace(0, t.num_rows, n_chunks + 1).astype(int)
    return pa.concat_tables([t.slice(a, b - a) for a, b in zip(cuts[:-1], cuts[1:])])


def test_column_chunks_walk_arrow_buffers_without_copy():
    log = make_log("tiny")
    t = _table(log, 4)
    for name in ("timestamp", "user_idx", "item_idx", "relevance"):
        chunks = list(mdp._column_chunks(t, name))
        assert len(chunks) == 4 and sum(c[2] for c in chunks) == len(log)
        for (addr, dt, cnt, keep), arrow_chunk in zip(chunks, t.column(name).chunks):
            assert addr == arrow_chunk.buffers()[1].address + arrow_chunk.offset * dt.itemsize      # the Arrow buffer itself
            got = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_char * (cnt * dt.itemsize)).from_address(addr)).view(dt)
            assert np.array_equal(got, arrow_chunk.to_numpy())
    # timestamp[ms] columns are viewed as int64 (order preserving), float32 / int64 ids are passed in their own dtype
    t2 = t.set_column(t.schema.get_field_index("timestamp"), "timestamp",
                      pa.chunked_array([pa.array(log["timestamp"].to_numpy() * 1000, type=pa.timestamp("ms"))]))
    (addr, dt, cnt, _), = mdp._column_chunks(t2, "timestamp")
    assert dt == np.int64 and cnt == len(log)
    t3 = t.set_column(t.schema.get_field_index("relevance"), "relevance", pa.chunked_array([pa.array(log["relevance"].astype(np.float32))]))
    assert next(mdp._column_chunks(t3, "relevance"))[1] == np.float32
    with pytest.raises(ValueError):
        list(mdp._column_chunks(pa.table({"user_idx": pa.array([1, None, 3])}), "user_idx"))
    # pandas: one chunk per column, the column's own array
    (addr, dt, cnt, keep), = mdp._column_chunks(log, "item_idx")
    assert cnt == len(log) and dt == log["item_idx"].dtype


def test_user_shard_bounds_balance_and_cover():
    rng = np.random.default_rng(0)
    counts = rng.integers(1, 400, size=5000)
    counts[17] = 60_000                                   # one heavy user
    for world in (1, 2, 3, 4, 8):
        b = mdp.user_shard_bounds(counts, world)
        assert b[0] == 0 and b[-1] == counts.size and np.all(np.diff(b) >= 0) and b.size == world + 1
        rows = np.array([counts[b[r]:b[r + 1]].sum() for r in range(world)])
        assert rows.sum() == counts.sum()
        assert rows.max() - counts.sum() / world <= counts.max()      # off by at most one user's episode
    log = make_log("tiny")
    for flavour in (log, pa.Table.from_pandas(log, preserve_index=False)):
        parts = [mdp.shard_log_by_user(flavour, r, 3) for r in range(3)]
        frames = [p.to_pandas() if isinstance(p, pa.Table) else p for p in parts]
        assert sum(len(f) for f in frames) == len(log)
        users = [set(f["user_idx"]) for f in frames]
        assert not (users[0] & users[1]) and not (users[1] & users[2]) and not (users[0] & users[2])
        assert max(users[0]) < min(users[1]) and max(users[1]) < min(users[2])      # contiguous ranges


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _split_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from replay_cql_b200.parallel import dist_info
        r, w, _ = dist_info()
        log = make_log("tiny")
        mine = mdp.shard_log_by_user(log, r, w)
        host = mdp.build_mdp(mine, top_k=3, seed=1)       # what this rank's GPU builder would be given
        np.savez(os.path.join(out_dir, f"shard{rank}.npz"), users=np.unique(mine["user_idx"]), n=len(mine),
                 obs=host.obs, term=host.term, rew=host.rew)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_user_split_world2_gloo(tmp_path):
    """Two ranks split the MDP build by user: disjoint contiguous user ranges, every row exactly once, and each shard's
    episodes equal the corresponding episodes of the unsharded build (rewards / terminals are per-user quantities)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    log = make_log("tiny")
    full = mdp.build_mdp(log, top_k=3, seed=1)
    z = [np.load(tmp_path / f"shard{r}.npz") for r in range(2)]
    assert z[0]["n"] + z[1]["n"] == len(log)
    assert z[0]["users"].max() < z[1]["users"].min()
    obs = np.concatenate([z[0]["obs"], z[1]["obs"]])
    assert np.array_equal(obs, full.obs)                  # user-major order: shard 0's episodes then shard 1's
    assert np.array_equal(np.concatenate([z[0]["term"], z[1]["term"]]), full.term)
    assert np.array_equal(np.concatenate([z[0]["rew"], z[1]["rew"]]), full.rew)


@pytest.mark.gpu
@pytest.mark.parametrize("flavour", ["arrow_chunks", "parquet_row_groups", "odd_dtypes", "pandas"])
def test_ingest_matches_host_builder_bit_exact(engine_factory, tmp_path, flavour):
    log = make_log("tiny").sample(frac=1.0, random_state=3).reset_index(drop=True)      # input order != episode order
    rng = np.random.default_rng(5)
    noise = rng.standard_normal(len(log)) * 1e-3
    ref = mdp.build_mdp(log, top_k=3, action_noise=noise)
    if flavour == "arrow_chunks":
        src = _table(log, 5)
    elif flavour == "parquet_row_groups":
        path = tmp_path / "log.parquet"
        pq.write_table(pa.Table.from_pandas(log, preserve_index=False), path, row_group_size=700)
        src = pq.read_table(path)
        assert src.column("user_idx").num_chunks > 1
    elif flavour == "odd_dtypes":
        src = pa.table({"user_idx": pa.array(log["user_idx"].to_numpy().astype(np.int64)),
                        "item_idx": pa.array(log["item_idx"].to_numpy().astype(np.int16)),
                        "timestamp": pa.array(log["timestamp"].to_numpy() * 1_000_000, type=pa.timestamp("us")),
                        "relevance": pa.array(log["relevance"].to_numpy().astype(np.float32))})
    else:
        src = log
    eng = engine_factory(batch_size=64)
    obs, act, rew, term, order = mdp.ingest_log(eng, src, top_k=3, action_noise=noise, want_outputs=True)
    assert eng.n_transitions == len(log)
    assert np.array_equal(obs, ref.obs) and np.array_equal(rew, ref.rew) and np.array_equal(term, ref.term)
    assert np.array_equal(act, ref.act) and np.array_equal(order, ref.order)
    rows = eng.export_transitions()
    tr = mdp.to_transitions(ref)
    want = np.concatenate([tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["term"], np.zeros((len(log), 1), np.float32)], 1)
    assert np.array_equal(rows, want)


@pytest.mark.gpu
def test_fit_from_arrow_table_equals_fit_from_pandas():
    """`CQL.fit(pyarrow.Table)` never converts the log to pandas and lands on the same model bit for bit."""
    from replay_cql_b200.models import CQL
    log = make_log("tiny")
    kw = dict(top_k=3, n_epochs=1, n_steps_per_epoch=12, batch_size=64, seed=3)
    a, b = CQL(**kw), CQL(**kw)
    a.fit(log)
    b.fit(_table(log, 4))
    assert np.array_equal(a.engine.get_state(), b.engine.get_state())
    assert a.fit_users.equals(b.fit_users) and a.fit_items.equals(b.fit_items)
    a.engine.close(); b.engine.close()
