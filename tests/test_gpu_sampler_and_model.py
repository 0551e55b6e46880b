"""GPU: replay table + sampler (K1) properties, and the CQL model end to end through the
Recommender API (fit / predict / predict_pairs / save / load), re-targeting the reference's
conformance tests (tests/models/test_all_models.py, test_save_load_models.py:48-70)."""
import numpy as np
import pandas as pd
import pytest
import torch

from replay_cql_b200.mdp import build_mdp, to_transitions
from replay_cql_b200.models import CQL
from replay_cql_b200.model_handler import load, save
from replay_cql_b200.synthetic import make_log

pytestmark = pytest.mark.gpu


def test_replay_table_and_sampler(engine_factory):
    log = make_log("tiny")
    mdp = build_mdp(log, top_k=10, seed=1)
    tr = to_transitions(mdp)
    N = len(mdp)
    eng = engine_factory(batch_size=64)
    eng.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
    assert eng.n_transitions == N
    # explicit indices: bit-exact rows (table build on the GPU == host expansion)
    idx = np.random.default_rng(0).integers(0, N, 777)
    rows = eng.sample_rows(777, idx=idx).cpu().numpy()
    want = np.concatenate([tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["term"], np.zeros((N, 1), np.float32)], 1)[idx]
    assert np.array_equal(rows, want)
    # one epoch of the permutation stream visits n_eff distinct rows exactly once
    n_eff = N - N % 64
    ep0 = eng.sample_rows(n_eff, pos=0).cpu().numpy()
    full = np.concatenate([tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["term"], np.zeros((N, 1), np.float32)], 1)
    # act carries per-row Gaussian noise -> rows are unique; map back to indices
    lookup = {full[i].tobytes(): i for i in range(N)}
    seen_idx = np.array([lookup[r.tobytes()] for r in ep0])
    assert len(set(seen_idx.tolist())) == n_eff
    ep1 = eng.sample_rows(n_eff, pos=n_eff).cpu().numpy()
    idx1 = np.array([lookup[r.tobytes()] for r in ep1])
    assert len(set(idx1.tolist())) == n_eff and not np.array_equal(idx1, seen_idx)   # fresh permutation
    assert abs(np.corrcoef(seen_idx[:-1], seen_idx[1:])[0, 1]) < 0.1               # shuffled, not a scan


def test_sampled_updates_run_and_are_deterministic(engine_factory):
    log = make_log("tiny")
    mdp = build_mdp(log, seed=1)
    outs = []
    for _ in range(2):
        eng = engine_factory(batch_size=64, seed=5)
        eng.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
        m = eng.update(20)
        assert all(np.isfinite(v) for v in m.values())
        outs.append(eng.get_state())
    assert np.array_equal(outs[0], outs[1])          # same seed -> bit-identical weights
    m_, v_, step = eng.get_optimizer()
    assert step == 20


@pytest.fixture(scope="module")
def fitted():
    log = make_log("tiny")
    model = CQL(top_k=3, n_epochs=2, batch_size=64, seed=3)
    model.fit(log)
    yield model, log
    model.engine.close()


def test_fit_predict_contract(fitted):
    model, log = fitted
    assert model._user_dim == 64 and model._item_dim == int(log["item_idx"].max()) + 1
    recs = model.predict(log, k=5)
    assert list(recs.columns) == ["user_idx", "item_idx", "relevance"]
    assert recs.groupby("user_idx").size().max() <= 5 and recs["user_idx"].nunique() == 64
    seen = set(zip(log["user_idx"], log["item_idx"]))
    assert not any((u, i) in seen for u, i in zip(recs["user_idx"], recs["item_idx"]))
    # subset of users / items, no filtering
    r2 = model.predict(log, k=3, users=[1, 2, 999], items=[0, 1, 2, 3], filter_seen_items=False)
    assert set(r2["user_idx"]) == {1, 2} and set(r2["item_idx"]) <= {0, 1, 2, 3}
    assert r2.groupby("user_idx").size().tolist() == [3, 3]


def test_predict_and_predict_pairs_agree(fitted):   # tests/models/test_all_models.py:63-96
    model, log = fitted
    recs = model.predict(log, k=4, filter_seen_items=False)
    pairs = model.predict_pairs(recs[["user_idx", "item_idx"]], log)
    merged = recs.merge(pairs, on=["user_idx", "item_idx"], suffixes=("_a", "_b"))
    assert len(merged) == len(recs)
    scale = max(1.0, merged["relevance_a"].abs().max())
    assert (merged["relevance_a"] - merged["relevance_b"]).abs().max() <= 1e-5 * scale
    assert model.predict_pairs(recs[["user_idx", "item_idx"]], log, k=2).groupby("user_idx").size().max() <= 2


def test_predict_k_above_kernel_limit(fitted):
    """The reference's predict takes any k (base_rec.py:466-539); above the kernel's per-pass limit (1024) the model
    selects in passes, each treating the items already picked as seen: same list as ranking every pair score."""
    from replay_cql_b200 import _lib
    model, log = fitted
    users = np.array([0, 5, 9], dtype=np.int32)
    items = np.arange(2500, dtype=np.int32)
    k = _lib.MAX_TOPK + 76
    from replay_cql_b200.mdp import seen_csr
    indptr, seen = seen_csr(log[log["user_idx"].isin(users)], 64)
    ti, ts = model._score_topk_any_k(users, items, k, indptr, seen, 64)
    assert ti.shape == (3, k)
    for r, u in enumerate(users):
        sc = model.engine.score_pairs(np.full(items.size, u, dtype=np.int32), items, mode="q")
        sc[seen[indptr[u]:indptr[u + 1]]] = -np.inf
        order = np.argsort(-sc, kind="stable")[:k]
        assert len(set(ti[r].tolist())) == k and not np.isin(ti[r], seen[indptr[u]:indptr[u + 1]]).any()
        assert np.all(np.diff(ts[r]) <= 0)
        np.testing.assert_allclose(ts[r], sc[order], rtol=1e-5, atol=1e-5 * np.abs(sc[order]).max())
        assert len(set(ti[r].tolist()) ^ set(order.tolist())) <= 4        # swaps among (near-)ties at the cut only


def test_save_load_roundtrip(fitted, tmp_path):     # tests/models/test_save_load_models.py:48-70
    model, log = fitted
    base = model.predict(log, k=5)
    path = str(tmp_path / "cql_model")
    save(model, path)
    loaded = load(path)
    assert str(loaded) == "CQL" and loaded._init_args == model._init_args
    again = loaded.predict(log, k=5)
    pd.testing.assert_frame_equal(base, again)
    loaded.engine.close()


def test_evaluate_matches_metrics_oracle(fitted):
    """CQL.evaluate (frames -> CSR -> cql_rank_metrics) == oracle on the model's own recommendations."""
    from oracle import metrics_oracle as M
    model, log = fitted
    recs = model.predict(log, k=5, filter_seen_items=False)
    gt = log.groupby("user_idx").tail(3)[["user_idx", "item_idx"]]
    got = model.evaluate(recs, gt, [1, 3, 5])
    recs_d = {int(u): g.sort_values(["relevance", "item_idx"], ascending=[False, True])["item_idx"].tolist()
              for u, g in recs.groupby("user_idx")}
    gt_d = {int(u): g["item_idx"].tolist() for u, g in gt.groupby("user_idx")}
    ref = M.rank_metrics(recs_d, gt_d, [1, 3, 5])
    for name in ref:
        for k in (1, 3, 5):
            assert got[name][k] == pytest.approx(ref[name][k], rel=1e-12, abs=1e-15), (name, k)


def test_score_policy_mode(fitted):
    model, log = fitted
    model.score = "policy"
    try:
        recs = model.predict(log, k=2)
        assert recs["relevance"].abs().max() <= 1.0     # tanh(mu)
    finally:
        model.score = "q"


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "f16x3"])
def test_phase_split_equals_fused_update(engine_factory, precision):
    """The four data-parallel phases (host all-reduce points in between; identity here) take exactly the
    same step as the fused CUDA-graph update: bit-identical weights after 5 updates."""
    log = make_log("tiny")
    mdp = build_mdp(log, seed=1)
    a = engine_factory(batch_size=64, seed=9, precision=precision)
    b = engine_factory(batch_size=64, seed=9, precision=precision)
    for e in (a, b):
        e.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
    a.update(5)
    for _ in range(5):
        b.update_data_parallel(lambda buf: None)
    torch.cuda.synchronize()
    assert np.array_equal(a.get_state(), b.get_state())
    assert a.get_optimizer()[2] == b.get_optimizer()[2] == 5


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_update_batches_equals_single_calls(engine_factory, precision):
    """cql_update_batches (pinned ring, pipelined copies, one graph per step) takes exactly the steps of repeated
    cql_update_batch calls: bit-identical state and per-step metrics, also when the ring wraps (n > 8)."""
    from tests import helpers as Hp
    B, n = 64, 11
    a = engine_factory(batch_size=B, seed=3, precision=precision)
    b = engine_factory(batch_size=B, seed=3, precision=precision)
    batches = [Hp.batch_to_numpy(Hp.make_batch(B, seed=300 + i, scale=1e-3)) for i in range(n)]
    single = [a.update_batch(bt)[0] for bt in batches]
    multi = b.update_batches(batches)
    torch.cuda.synchronize()
    assert np.array_equal(a.get_state(), b.get_state())
    for s1, s2 in zip(single, multi):
        assert s1 == s2
    assert b.update_batches([]) == []


@pytest.mark.parametrize("top_k", [0, 1, 3, 1000])
def test_gpu_mdp_builder_bit_exact(engine_factory, top_k):
    """cql_build_mdp (stable radix sorts on the device) == host builder == loop oracle, bit for bit,
    including timestamp ties, shuffled rows and duplicate (user, item) rows."""
    from oracle import mdp_oracle
    from replay_cql_b200.mdp import build_mdp_on_device
    rng = np.random.default_rng(11)
    n = 5000
    log = pd.DataFrame({"user_idx": rng.integers(0, 97, n), "item_idx": rng.integers(0, 300, n),
                        "timestamp": rng.integers(0, 50, n), "relevance": rng.integers(1, 6, n).astype(float)})
    noise = rng.standard_normal(n) * 1e-3
    eng = engine_factory(batch_size=64)
    obs, act, rew, term, order = build_mdp_on_device(eng, log, top_k=top_k, action_noise=noise, want_outputs=True)
    host = build_mdp(log, top_k=top_k, action_noise=noise)
    assert np.array_equal(order, host.order)
    assert np.array_equal(obs, host.obs) and np.array_equal(act, host.act)
    assert np.array_equal(rew, host.rew) and np.array_equal(term, host.term)
    ref = mdp_oracle.build_mdp(log, top_k=top_k, action_noise=noise)
    assert np.array_equal(obs, ref["obs"]) and np.array_equal(rew, ref["rew"][:, 0]) and np.array_equal(term, ref["term"][:, 0])
    # the table itself: rows by explicit index == host expansion
    tr = to_transitions(host)
    full = np.concatenate([tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["term"], np.zeros((n, 1), np.float32)], 1)
    idx = rng.integers(0, n, 500)
    assert np.array_equal(eng.sample_rows(500, idx=idx).cpu().numpy(), full[idx])


def test_gpu_mdp_builder_datetime_and_seeded_noise(engine_factory):
    from replay_cql_b200.mdp import build_mdp_on_device
    log = make_log("tiny").sample(frac=1.0, random_state=5).reset_index(drop=True)
    log["timestamp"] = pd.to_datetime(log["timestamp"], unit="s")
    eng = engine_factory(batch_size=64, seed=3)
    obs, act, rew, term, order = build_mdp_on_device(eng, log, top_k=10, want_outputs=True)
    host = build_mdp(log, top_k=10, action_noise=np.zeros(len(log)))
    assert np.array_equal(order, host.order) and np.array_equal(rew, host.rew) and np.array_equal(term, host.term)
    d = act - host.act                       # seeded N(0, 1e-3) draw on the device
    assert 5e-4 < d.std() < 2e-3 and abs(d.mean()) < 2e-4


def test_data_parallel_stepper_takes_exactly_n_steps(engine_factory):
    """the stepper's eager warm-up updates are REAL updates and must come out of the requested budget
    (the N>1 fit used to take total + 3 steps).  World 1 with a no-op gradient exchange: the 4 phases + graph capture
    are the product path; the result must be bit-identical to the single-GPU `cql_update` loop of the same length."""
    from replay_cql_b200.parallel import DataParallelStepper
    log = make_log("tiny")
    mdp = build_mdp(log, seed=1)
    ref = engine_factory(batch_size=64, seed=5, precision="f16x3")
    ref.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
    ref.update(17)
    for first, second in ((17, 0), (2, 15), (5, 12)):
        eng = engine_factory(batch_size=64, seed=5, precision="f16x3")
        eng.load_transitions(mdp.obs, mdp.act, mdp.rew, mdp.term)
        stepper = DataParallelStepper(eng, reducer=lambda buffer_id: None)
        stepper.run(first)
        stepper.run(second)
        stepper.finish()
        assert stepper.steps_done == 17 and stepper.eager_steps == 3
        assert eng.get_optimizer()[2] == 17
        assert np.array_equal(eng.get_state(), ref.get_state())


@pytest.mark.parametrize("with_table", [False, True])
def test_exact_resume_after_save_load(tmp_path, with_table, caplog):
    """SURVEY.md 8f-4 / VERDICT r01 item 7: train N -> save -> load -> train M is bit-identical to train N + M:
    weights, targets, Adam moments, step counter (= position in the epoch permutation and Philox counter).  The replay
    table either travels in the checkpoint (``save_replay_table=True``) or is rebuilt from the log (deterministic MDP
    builder).  Checkpoint layout = the reference's (``replay/model_handler.py:29-53``)."""
    import logging
    log = make_log("tiny")
    N, M = 37, 23
    kw = dict(top_k=3, batch_size=64, seed=11, n_epochs=1, save_replay_table=with_table, log_every=10)
    full = CQL(n_steps_per_epoch=N + M, **kw)
    with caplog.at_level(logging.DEBUG, logger="replay"):
        full.fit(log)
    # the six losses are logged every `log_every` steps (d3rlpy's progress line, the house logger)
    lines = [r.getMessage() for r in caplog.records if "critic_loss=" in r.getMessage()]
    assert len(lines) == 6 and all(k in lines[-1] for k in ("temp_loss=", "temp=", "alpha_loss=", "alpha=", "actor_loss="))
    assert f"CQL step {N + M} " in lines[-1]
    part = CQL(n_steps_per_epoch=N, **kw)
    part.fit(log)
    path = str(tmp_path / "ckpt")
    save(part, path)
    part.engine.close()
    resumed = load(path)
    assert resumed.engine.get_optimizer()[2] == N
    assert resumed.engine.n_transitions == (full.engine.n_transitions if with_table else 0)
    if with_table:
        assert np.array_equal(resumed.engine.export_transitions(), full.engine.export_transitions())
        resumed.resume(n_steps=M)
    else:
        with pytest.raises(ValueError):
            resumed.resume(n_steps=M)
        resumed.resume(log, n_steps=M)
    assert np.array_equal(resumed.engine.get_state(), full.engine.get_state())
    m1, v1, s1 = resumed.engine.get_optimizer()
    m2, v2, s2 = full.engine.get_optimizer()
    assert s1 == s2 == N + M and np.array_equal(m1, m2) and np.array_equal(v1, v2)
    pd.testing.assert_frame_equal(resumed.predict(log, k=5), full.predict(log, k=5))
    full.engine.close()
    resumed.engine.close()
