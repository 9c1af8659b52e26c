"""Multi-GPU DSGD path. CPU: partitioning / strip extraction and the N>1 host plumbing over a
world_size-2 gloo group. GPU: several logical ranks on ONE device (one context + one host thread
per rank, peer pointers inside the process) against the single-GPU session and the oracle."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(U=1200, I=300, n=40000, seed=21):
    tr, te = cu.synth_ratings(U, I, n, rank=4, noise=0.3, integer_ratings=True, seed=seed)
    return tr, te, U, I


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_and_strips_cover_the_problem_exactly(world):
    tr, te, U, I = _problem()
    part = cu.dsgd_partition(tr, U, I, world)
    assert part.users_per_block.sum() == U and part.item_block_ptr[-1] == I
    assert sorted(part.item_new.tolist()) == list(range(I))
    # balanced by rating count (LPT): every block row / column within 2 % of the mean
    assert part.block_nnz.sum() == len(tr)
    for marg in (part.block_nnz.sum(0), part.block_nnz.sum(1)):
        assert marg.max() <= 1.02 * marg.mean() + 1
    got_train = []
    for r in range(world):
        strip = cu.dsgd_extract_strip(tr, part, r)
        assert np.all(np.diff(strip["user"]) >= 0) and strip["user"].max() < part.users_per_block[r]
        users = np.flatnonzero(part.user_block == r)
        inv = np.empty(I, np.int64)
        inv[part.item_new] = np.arange(I)
        back = np.zeros(len(strip), dtype=cu.RATING_DTYPE)
        back["user"], back["item"], back["rating"] = users[strip["user"]], inv[strip["item"]], strip["rating"]
        got_train.append(back)
        # per-user order is preserved (the sampler relies on it)
        u0 = users[0]
        assert np.array_equal(back[back["user"] == u0], tr[tr["user"] == u0])
        # item blocks are contiguous ranges in the renumbered space
        blk = np.searchsorted(part.item_block_ptr, strip["item"], side="right") - 1
        assert np.array_equal(np.bincount(blk, minlength=world), part.block_nnz[r])
    allr = np.sort(np.concatenate(got_train), order=["user", "item"])
    assert np.array_equal(allr, np.sort(tr, order=["user", "item"]))


def test_item_keep_fractions_bound_every_items_load():
    """Host part of the opt-in item-step thinning: lr x share x groups x keep <= budget for every item,
    keep = 1 below the budget, shares taken per item block under per-user sampling."""
    tr, te, U, I = _problem(U=3000, I=400, n=120000)
    world, lr, groups, budget = 4, 0.01, 768, 0.1
    part = cu.dsgd_partition(tr, U, I, world)
    for r in range(world):
        strip = cu.createSparseMatrix(cu.dsgd_extract_strip(tr, part, r), int(part.users_per_block[r]), I)
        keep = cu.dsgd_item_keep(strip, part.item_block_ptr, lr, groups, budget)
        deg = np.diff(strip.indptr)
        w = np.bincount(strip.indices, weights=np.repeat(1.0 / np.maximum(deg, 1), deg), minlength=I)
        blk = np.searchsorted(part.item_block_ptr, np.arange(I), side="right") - 1
        share = w / np.bincount(blk, weights=w, minlength=world)[blk]
        load = lr * share * groups
        want = np.where(load > budget, budget / np.maximum(load, 1e-300), 1.0)
        assert np.allclose(keep, want, rtol=1e-5) and keep.max() == 1.0 and 0 < keep.min() < 0.2
        assert np.all(load * keep <= budget * (1 + 1e-5))
        assert (keep < 1).sum() > 20  # measured on B200 with these parameters: 62-68 items per rank
    one = cu.dsgd_item_keep(cu.createSparseMatrix(tr, U, I), None, lr, groups, 1e9)
    assert np.all(one == 1.0)
    with pytest.raises(cu._lib.Cu2bError):
        cu.dsgd_item_keep(cu.createSparseMatrix(tr, U, I), np.array([0, 5, I - 1], np.int32), lr, groups, 0.5)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import cu2rec_b200 as cu, oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
tr, te = cu.synth_ratings(600, 200, 15000, rank=4, noise=0.3, seed=5)
U, I, k = 600, 200, 8
part = cu.dsgd_partition(tr, U, I, world)          # every rank computes the same partition
chk = torch.tensor([int(part.item_new.astype(np.int64).sum() * 7 + part.user_block.astype(np.int64).dot(np.arange(U)))])
both = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
assert both[0].item() == both[1].item(), "partition differs between ranks"
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
inp = cu.dsgd_rank_inputs(tr, te, U, I, part, rank, P, Q, ub, ib)
mu = np.float32(tr["rating"].astype(np.float64).mean())
# rank-local loss sums on the strip (CPU oracle) + all-reduce == loss of the whole problem
_, _, sse, sae = O.loss(inp.train.indptr, inp.train.indices, inp.train.data, inp.P, inp.Q, inp.user_bias, inp.item_bias, mu, k)
t = torch.tensor([sse, sae, float(inp.train.nonzeros)], dtype=torch.float64)
dist.all_reduce(t)
full = cu.createSparseMatrix(tr, U, I)
_, _, gsse, gsae = O.loss(full.indptr, full.indices, full.data, P, Q, ub, ib, mu, k)
assert abs(t[0].item() - gsse) / gsse < 1e-12 and abs(t[1].item() - gsae) / gsae < 1e-12 and int(t[2].item()) == len(tr)
# the handle exchange used by bench.py: fixed-size blobs all-gathered in rank order
blob = bytes([rank + 1]) * cu._lib.DSGD_HANDLE_BYTES
g = [torch.zeros(cu._lib.DSGD_HANDLE_BYTES, dtype=torch.uint8) for _ in range(world)]
dist.all_gather(g, torch.frombuffer(bytearray(blob), dtype=torch.uint8))
assert [bytes(x.numpy().tobytes())[0] for x in g] == [1, 2]
# the per-user sampler is keyed by ORIGINAL user ids: the union of both ranks' draws is the single-GPU draw
s_local = O.sample_per_user(inp.train.indptr, inp.train.indices, inp.train.data, 42, 0, 1)
print("OK", rank, len(s_local))
dist.destroy_process_group()
'''


def test_world2_gloo_host_logic(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, e[-2000:]


def _run_logical_ranks(world, tr, te, U, I, k, iters, ce, device=0):
    part = cu.dsgd_partition(tr, U, I, world)
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    ranks = []
    for r in range(world):
        inp = cu.dsgd_rank_inputs(tr, te, U, I, part, r, P, Q, ub, ib)
        cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=ce)
        ranks.append(cu.Dsgd(r, world, inp, part, cfg, mu, device=device))
    handles = [d.handle for d in ranks]
    for d in ranks:
        d.connect(handles)
    errs = []

    def work(d):
        try:
            d.run(iters)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(d,)) for d in ranks]
    [t.start() for t in th]
    [t.join(timeout=240) for t in th]
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "DSGD ranks did not finish"
    return part, ranks, (P, Q, ub, ib, mu)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
def test_dsgd_logical_ranks_match_single_gpu_training(world):
    tr, te, U, I = _problem(U=3000, I=400, n=120000)
    k, iters, ce = 16, 120, 40
    part, ranks, (P, Q, ub, ib, mu) = _run_logical_ranks(world, tr, te, U, I, k, iters, ce)
    logs = [d.log() for d in ranks]
    for lg in logs[1:]:
        assert lg == logs[0]  # every rank combines the partial sums in rank order: identical bits
    assert [r["iteration"] for r in logs[0]] == [1, 40, 80, 120]
    assert sum(d.stats()["updates"] for d in ranks) == iters * U
    # single-GPU session on the same problem / init / sampler stream
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=ce)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(iters)
        ref = s.log()
    for a, b in zip(logs[0], ref):
        for key in ("train_rmse", "test_rmse"):
            assert abs(a[key] - b[key]) / b[key] < 0.01, (key, a, b)
    assert abs(logs[0][-1]["test_rmse"] - ref[-1]["test_rmse"]) / ref[-1]["test_rmse"] < 0.005
    # the model gathered from the ranks reproduces the logged loss (oracle, whole problem)
    Pg, ubg = np.empty((U, k), np.float32), np.empty(U, np.float32)
    inv = np.empty(I, np.int64)
    inv[part.item_new] = np.arange(I)
    Qg = ibg = None
    for r, d in enumerate(ranks):
        Ps, Qn, ubs, ibn = d.download()
        users = np.flatnonzero(part.user_block == r)
        Pg[users], ubg[users] = Ps, ubs
        if r == 0:
            Qg, ibg = Qn[part.item_new], ibn[part.item_new]  # back to original item order
    _, rmse, _, _ = O.loss(mte.indptr, mte.indices, mte.data, Pg, Qg, ubg, ibg, mu, k)
    assert abs(rmse - logs[0][-1]["test_rmse"]) / rmse < 1e-4
    for d in ranks:
        d.close()


@pytest.mark.gpu
def test_dsgd_local_sums_add_up_to_the_logged_loss():
    tr, te, U, I = _problem(U=1000, I=200, n=30000)
    part, ranks, (_, _, _, _, mu) = _run_logical_ranks(2, tr, te, U, I, 8, 20, 20)
    sums = np.array([d.local_sums() for d in ranks]).sum(0)  # what an NCCL / gloo all-reduce would produce
    last = ranks[0].log()[-1]
    assert abs(np.sqrt(sums[0] / len(tr)) - last["train_rmse"]) < 1e-6
    assert abs(sums[3] / len(te) - last["test_mae"]) < 1e-6
    for d in ranks:
        d.close()

@pytest.mark.gpu
def test_dsgd_reload_equals_fresh_contexts():
    """cu2b_dsgd_reload on two logical ranks: train, reload the same strips + initial model,
    train again -> the log and the downloaded strips of freshly created contexts, bit for bit,
    on a problem whose items are never shared between users (any schedule is deterministic)."""
    rng = np.random.RandomState(4)
    U, k, iters, ce, world = 600, 32, 48, 16, 2
    deg = rng.randint(1, 6, U)
    n = int(deg.sum())
    tr = np.zeros(n, dtype=cu.RATING_DTYPE)
    tr["user"], tr["item"], tr["rating"] = np.repeat(np.arange(U), deg), rng.permutation(n), rng.randint(1, 6, n)
    tr = tr[np.lexsort((tr["item"], tr["user"]))]
    te = tr[::3].copy()
    part, ranks, (P, Q, ub, ib, mu) = _run_logical_ranks(world, tr, te, U, n, k, iters, ce)
    fresh = [(d.log(), d.download()) for d in ranks]
    inputs = [cu.dsgd_rank_inputs(tr, te, U, n, part, r, P, Q, ub, ib) for r in range(world)]
    for d, inp in zip(ranks, inputs):
        d.reload(inp, mu)  # every rank reloads before any rank runs (the "barrier")
    th = [threading.Thread(target=d.run, args=(iters,)) for d in ranks]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not any(t.is_alive() for t in th)
    for d, (lg, model) in zip(ranks, fresh):
        assert d.log() == lg
        for a, b in zip(d.download(), model):
            assert np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
        d.close()


# ---------------------------------------------------------------------------------------------
# Opt-in item-step thinning (CU2B_DSGD_THIN; default off in the product). The model behind it is
# tools/async_sim; all three tests passed on B200 in round 1.
# ---------------------------------------------------------------------------------------------


def _with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.gpu
def test_dsgd_thinning_with_unreachable_budget_is_the_default_path_bit_for_bit():
    """keep = 1 for every item: the THIN kernel and the flagging sampler must reproduce the default
    kernels exactly (disjoint items => any schedule is deterministic)."""
    rng = np.random.RandomState(4)
    U, k, iters, ce, world = 600, 32, 48, 16, 2
    deg = rng.randint(1, 6, U)
    n = int(deg.sum())
    tr = np.zeros(n, dtype=cu.RATING_DTYPE)
    tr["user"], tr["item"], tr["rating"] = np.repeat(np.arange(U), deg), rng.permutation(n), rng.randint(1, 6, n)
    tr = tr[np.lexsort((tr["item"], tr["user"]))]
    te = tr[::3].copy()
    _, base, _ = _run_logical_ranks(world, tr, te, U, n, k, iters, ce)
    _, thin, _ = _with_env("CU2B_DSGD_THIN", "1e9", lambda: _run_logical_ranks(world, tr, te, U, n, k, iters, ce))
    for a, b in zip(base, thin):
        assert a.log() == b.log()
        for x, y in zip(a.download(), b.download()):
            assert np.array_equal(np.asarray(x).view(np.uint32), np.asarray(y).view(np.uint32))
        a.close()
        b.close()


@pytest.mark.gpu
def test_dsgd_thinning_trains_to_the_same_rmse():
    tr, te, U, I = _problem(U=3000, I=400, n=120000)
    k, iters, ce, world = 16, 480, 120, 4
    _, base, _ = _run_logical_ranks(world, tr, te, U, I, k, iters, ce)
    # 100 items per block: the hottest one receives ~18 % of its block's draws, so the default path caps the
    # user groups in flight at 0.5 / (lr * 0.18) = 284 of 768. Thinning with the same budget runs all 768 and
    # keeps 37 % of that item's item-side steps. Popular items learn more slowly at first (measured on B200
    # with a budget of 0.1, lowest keep 0.074: +2.8 % test RMSE after 160 iterations), so compare later.
    _, thin, _ = _with_env("CU2B_DSGD_THIN", "0.5", lambda: _run_logical_ranks(world, tr, te, U, I, k, iters, ce))
    a, b = base[0].log()[-1], thin[0].log()[-1]
    print("default", [round(r["test_rmse"], 4) for r in base[0].log()], "thinned", [round(r["test_rmse"], 4) for r in thin[0].log()])
    # measured on B200: default [1.0292, 0.9737, 0.8407, 0.6697, 0.5689], thinned [1.0291, 0.9764, 0.8567, 0.6816, 0.5750]
    # (still in the steep part of the descent after 480 iterations: +1.1 %)
    assert np.isfinite(b["test_rmse"]) and abs(a["test_rmse"] - b["test_rmse"]) / a["test_rmse"] < 0.025, (a, b)
    assert base[0].log() != thin[0].log()  # some item-side steps really were skipped
    for d in base + thin:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["rows_and_bias", "bias_only", "rows_0.0002_bias_0.0001"])
def test_dsgd_thinning_bit_exact_against_oracle_replay_on_disjoint_items(what):
    """With items never shared between users every schedule is deterministic, so the thinned DSGD run can be
    replayed on the CPU: a user's draws of a round grouped by item block in the rank's sub-epoch order, the
    item side frozen for the draws whose second Philox word falls above keep[item] (cu2b_dsgd_item_keep)."""
    import ctypes as C
    rng = np.random.RandomState(8)
    U, k, iters, ce, world, grid, budget, lr = 400, 32, 48, 16, 2, 2, 2e-4, 0.01
    deg = rng.randint(1, 6, U)
    n = int(deg.sum())
    tr = np.zeros(n, dtype=cu.RATING_DTYPE)
    tr["user"], tr["item"], tr["rating"] = np.repeat(np.arange(U), deg), rng.permutation(n), rng.randint(1, 6, n)
    tr = tr[np.lexsort((tr["item"], tr["user"]))]
    te = tr[::3].copy()

    # budgets of the row steps and of the bias steps (0 = that switch is not set); a draw that skips the row
    # step always skips the bias step too
    b_rows, b_bias = {"rows_and_bias": (budget, 0.0), "bias_only": (0.0, budget), "rows_0.0002_bias_0.0001": (budget, budget / 2)}[what]
    env = {"CU2B_DSGD_GRID": str(grid)}
    if b_rows:
        env["CU2B_DSGD_THIN"] = repr(b_rows)
    if b_bias:
        env["CU2B_DSGD_THIN_BIAS"] = repr(b_bias)

    def run(names=tuple(env)):
        if not names:
            return _run_logical_ranks(world, tr, te, U, n, k, iters, ce)
        return _with_env(names[0], env[names[0]], lambda: run(names[1:]))
    part, ranks, (P, Q, ub, ib, mu) = run()
    groups = grid * 8 * (32 // 8)  # kp = 32 -> 8 lanes per rating, 4 lane groups per warp, 8 warps per CTA
    keeps_row, keeps_bias = [], []
    for r in range(world):
        strip = cu.createSparseMatrix(cu.dsgd_extract_strip(tr, part, r), int(part.users_per_block[r]), n)
        one = np.ones(n, np.float32)
        kr = cu.dsgd_item_keep(strip, part.item_block_ptr, lr, groups, b_rows) if b_rows else one
        kb = np.minimum(kr, cu.dsgd_item_keep(strip, part.item_block_ptr, lr, groups, b_bias)) if b_bias else kr
        keeps_row.append(kr)
        keeps_bias.append(kb)
        assert 0.01 < kb.min() < 0.9  # thinning is really active
    # replay in original ids
    full = cu.createSparseMatrix(tr, U, n)
    item_blk = np.searchsorted(part.item_block_ptr, part.item_new, side="right") - 1
    Po, Qo, ubo, ibo = P.copy(), Q.copy(), ub.copy(), ib.copy()
    hyp = {}
    for code in (0, 1, 2, 3):  # oracle is_train codes: 0 none, 1 row + bias, 2 bias only, 3 row only
        hyp[code] = O.hyper(k)
        hyp[code].is_train = code
    bounds = [0, 1] + list(range(ce, iters + 1, ce))  # segments end after every check iteration (1, 16, 32, 48)
    out, key = (C.c_uint32 * 4)(), (C.c_uint32 * 2)(42, 0x43553242)
    frozen_total = 0
    for a, b in zip(bounds[:-1], bounds[1:]):
        draws = O.sample_per_user(full.indptr, full.indices, full.data, 42, a, b - a).reshape(b - a, U)
        stream, flags = [], []
        for u in range(U):
            r = int(part.user_block[u])
            for sub in range(world):
                blk = (r + sub) % world
                for t in range(b - a):
                    d = draws[t, u]
                    if item_blk[d["item"]] != blk:
                        continue
                    O.lib().orc_philox4x32_10((C.c_uint32 * 4)(u, a + t, 0, 0x53474431), key, out)
                    x = np.float32(out[1] >> 8) * np.float32(1.0 / 16777216.0)
                    new_id = part.item_new[d["item"]]
                    row_on, bias_on = x < keeps_row[r][new_id], x < keeps_bias[r][new_id]
                    stream.append(d)
                    flags.append({(True, True): 1, (False, True): 2, (True, False): 3, (False, False): 0}[(bool(row_on), bool(bias_on))])
        stream, flags = np.array(stream, dtype=O.TRIPLET), np.array(flags)
        frozen_total += int((flags != 1).sum())
        cuts = np.flatnonzero(np.diff(flags)) + 1
        for lo, hi in zip(np.r_[0, cuts], np.r_[cuts, len(stream)]):
            Po, Qo, ubo, ibo = O.sgd_apply_stream(stream[lo:hi], Po, Qo, ubo, ibo, mu, hyp[int(flags[lo])], O.FLAVOUR_KERNEL)
    assert frozen_total > iters * U // 10
    Qg = ibg = None
    for r, d in enumerate(ranks):
        Ps, Qn, ubs, ibn = d.download()
        users = np.flatnonzero(part.user_block == r)
        assert np.array_equal(Ps.view(np.uint32), Po.reshape(U, k)[users].view(np.uint32))
        assert np.array_equal(ubs.view(np.uint32), ubo[users].view(np.uint32))
        if r == 0:
            Qg, ibg = Qn[part.item_new], ibn[part.item_new]
        d.close()
    assert np.array_equal(Qg.view(np.uint32), Qo.reshape(n, k).view(np.uint32))
    assert np.array_equal(ibg.view(np.uint32), ibo.view(np.uint32))


# ---------------------------------------------------------------------------------------------
# DSGD behind the reference's own surface: cu2b_train() / train() with config token 17 (n_gpus) > 1
# (mf.cu:61 -> training.h:12-15). CU2B_LOGICAL_RANKS=1 lets the ranks share the one test device.
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_train_entry_point_runs_dsgd_when_n_gpus_is_set(world):
    tr, te, U, I = _problem(U=3000, I=400, n=120000)
    k, iters, ce = 16, 120, 40
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    one = cu.train(mtr, mte, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), mu)
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=ce, n_gpus=world)
    out = _with_env("CU2B_LOGICAL_RANKS", "1", lambda: cu.train(mtr, mte, cfg, mu))
    assert [r["iteration"] for r in out["log"]] == [1, 40, 80, 120]
    assert out["stats"]["updates"] == iters * U and cfg.cur_iterations == iters
    for a, b in zip(out["log"], one["log"]):
        for key in ("train_rmse", "test_rmse"):
            assert abs(a[key] - b[key]) / b[key] < 0.01, (key, a, b)
    assert abs(out["log"][-1]["test_rmse"] - one["log"][-1]["test_rmse"]) / one["log"][-1]["test_rmse"] < 0.005
    # the returned model is in the caller's ids: the oracle's loss on it is the logged loss
    _, rmse, _, _ = O.loss(mte.indptr, mte.indices, mte.data, out["P"], out["Q"], out["user_bias"], out["item_bias"], mu, k)
    assert abs(rmse - out["log"][-1]["test_rmse"]) / rmse < 1e-4
    assert np.array_equal(np.isnan(out["losses"]), np.isnan(one["losses"]))


def test_train_entry_point_rejects_more_gpus_than_the_box_has():
    tr, te, U, I = _problem(U=200, I=50, n=3000)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    cfg = cu.Config(total_iterations=4, n_factors=8, check_error=2, n_gpus=64)
    with pytest.raises(cu._lib.Cu2bError):
        cu.train(mtr, mte, cfg, 3.5)


def test_partition_places_one_popular_row_per_l2_block():
    """Row placement inside an item block (cu2b_paired_slots): the two 512-byte rows of a 1 KB block share one
    pair of L2 slices, so the popular half of a block's items sits on the even matrix rows in popularity order and the
    other half on the odd rows -- every slice pair holds exactly one popular row, block weights fall along the range."""
    tr, te, U, I = _problem(U=3000, I=401, n=120000)
    counts = np.bincount(tr["item"], minlength=I)
    for world in (1, 3, 4):
        part = cu.dsgd_partition(tr, U, I, world)
        inv = np.empty(I, np.int64)
        inv[part.item_new] = np.arange(I)          # matrix row -> original item
        for b in range(world):
            r0, r1 = int(part.item_block_ptr[b]), int(part.item_block_ptr[b + 1])
            rows = np.arange(r0, r1)
            c = counts[inv[rows]]
            even, odd = c[rows % 2 == 0], c[rows % 2 == 1]
            assert np.all(np.diff(even) <= 0), "popular rows in popularity order on the even rows"
            assert np.all(np.diff(odd) <= 0)
            assert even.min() >= odd.max(), "every even row is at least as popular as every odd row"


def test_partition_and_strips_are_the_same_serial_and_threaded(monkeypatch):
    """cu2b_dsgd_partition / cu2b_dsgd_extract_strip cut large inputs into per-thread slices (per-thread histograms,
    count-then-write): same partition, same block counts, same strips as the single-threaded pass, also when a slice
    boundary falls inside a user's ratings and when an id is out of range."""
    tr, _ = cu.synth_ratings(1500, 333, 60000, rank=4, noise=0.3, seed=8)
    U, I = 1500, 333
    outs = []
    for min_bytes in ("0", str(1 << 40)):
        monkeypatch.setenv("CU2B_IO_PARALLEL_MIN_BYTES", min_bytes)
        for threads in (("3", "8") if min_bytes == "0" else ("1",)):
            monkeypatch.setenv("OMP_NUM_THREADS", threads)
            part = cu.dsgd_partition(tr, U, I, 4)
            strips = [cu.dsgd_extract_strip(tr, part, r) for r in range(4)]
            outs.append((part, strips))
    ref_part, ref_strips = outs[-1]
    assert sum(len(s) for s in ref_strips) == len(tr) and int(ref_part.block_nnz.sum()) == len(tr)
    for part, strips in outs[:-1]:
        for name in ("user_block", "user_local", "users_per_block", "item_new", "item_block_ptr", "block_nnz"):
            assert np.array_equal(getattr(part, name), getattr(ref_part, name)), name
        for a, b in zip(strips, ref_strips):
            assert a.tobytes() == b.tobytes()
    monkeypatch.setenv("CU2B_IO_PARALLEL_MIN_BYTES", "0")
    bad = tr.copy()
    bad["item"][41234] = I
    with pytest.raises(cu._lib.Cu2bError) as err:
        cu.dsgd_partition(bad, U, I, 4)
    assert "41234" in str(err.value)
