"""Row-sharded mode (BASELINE configs[4]).  On a single GPU two shards are driven as two handles in two
host threads with a host-side rank-ordered exchange (no kernel ever waits on another launch); with >= 2 GPUs
the in-library NCCL exchange is exercised across two processes."""
import os
import socket
import threading
import numpy as np
import pytest
import oracle
from helpers import synth, PRIOR_CASES
from mcmcglm_b200 import Engine
from mcmcglm_b200.multigpu import shard_rows, ordered_sum, DeviceBuffer

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _run_two_shards_one_gpu(family, prior, X, y, beta0, w, iters, replay_u=None, seed=0, sd=1.0, **ekw):
    import torch
    world = 2
    n, p = X.shape
    bar = threading.Barrier(world)
    slots = [None] * world
    out = [None] * world
    errs = []

    def rank_main(r):
        try:
            lo, hi = shard_rows(n, world, r)
            e = Engine(hi - lo, p, family=family, sd=sd, w=w, n_chains=1, K=6, driver="stepwise", row_sharded=True,
                       seed=seed, **PRIOR_CASES[prior], **ekw)

            def xfn(ptr, count, stream):
                t = torch.as_tensor(DeviceBuffer(ptr, count), device="cuda")
                torch.cuda.synchronize()
                slots[r] = t.cpu().numpy().copy()
                bar.wait()
                tot = ordered_sum(slots)
                bar.wait()
                t.copy_(torch.from_numpy(tot))
                torch.cuda.synchronize()
                return 0
            e.set_exchange(xfn)
            e.set_data(X[lo:hi], y[lo:hi])
            e.init_chain(0, beta0)
            S, st = e.run(iters, replay_u=replay_u)
            f = e.log_potential(0, 1, [beta0[1] * 0 + 0.123])
            out[r] = (S[0], st, e.state(0), f, (lo, hi))
            e.close()
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)
            bar.abort()
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    return out


def test_readme_chain_row_sharded_in_two():
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    out = _run_two_shards_one_gpu("gaussian", "normal", z["X"], z["y"], z["beta0"], 0.5, 200, replay_u=z["uniforms"])
    for S, st, state, f, (lo, hi) in out:
        assert np.max(np.abs(S - z["samples"][1:201])) <= 1e-9           # every rank holds the reference's chain
        assert st["ref_evals"] > 0
        assert st["jet_passes"] >= st["updates"] - st["jet_fallbacks"] > 0   # one exchanged pass per update (jet passes)
    assert np.array_equal(out[0][0], out[1][0])                          # ranks took identical branches, bit for bit
    assert out[0][3][0] == out[1][3][0]
    eta_full = np.concatenate([out[0][2][1], out[1][2][1]])
    assert np.max(np.abs(eta_full - z["X"] @ out[0][2][0])) < 1e-11


@pytest.mark.parametrize("family,prior", [("binomial", "laplace"), ("poisson", "student_t")])
def test_sharded_matches_oracle(family, prior):
    X, y, bt = synth(family, 4003, 4, seed=12)
    m = oracle.make_model(family, **PRIOR_CASES[prior])
    b0 = np.zeros(4)
    ref = oracle.run_chain(m, X, y, b0, w=0.3, n_iter=25, seed=9, chain=0)
    out = _run_two_shards_one_gpu(family, prior, X, y, b0, 0.3, 25, seed=9)
    for S, st, *_ in out:
        assert np.max(np.abs(S - ref["samples"])) <= 1e-9
        assert st["jet_passes"] > 0
    # and bit for bit the chain of the exact passes (column statistics reduced over the shards, per-shard constants)
    exact = _run_two_shards_one_gpu(family, prior, X, y, b0, 0.3, 25, seed=9, jet=False)
    assert exact[0][1]["jet_passes"] == 0
    assert np.array_equal(out[0][0], exact[0][0]) and np.array_equal(out[1][0], exact[1][0])
    for k in ("ref_evals", "stepouts", "shrinks", "uniforms_used"):
        assert out[0][1][k] == exact[0][1][k]


def _run_two_shards_mailboxes(family, prior, X, y, beta0, w, iters, replay_u=None, seed=0, **ekw):
    """Two row shards as two handles on ONE device, both on the persistent driver: their sweep kernels run side by side
    (small grids) and exchange every pass's sums through each other's mailboxes -- the very code path of the multi-GPU
    run, with the peer pointers being plain device pointers of the same process instead of NVLink mappings."""
    import torch
    world = 2
    n, p = X.shape
    bar = threading.Barrier(world)
    slots, out, errs, eng, ptrs = [None] * world, [None] * world, [], [None] * world, [None] * world

    def rank_main(r):
        try:
            lo, hi = shard_rows(n, world, r)
            e = Engine(hi - lo, p, family=family, sd=1.0, w=w, n_chains=1, K=6, driver="grid", row_sharded=True,
                       seed=seed, **PRIOR_CASES[prior], **ekw)
            eng[r] = e

            def xfn(ptr, count, stream):          # only used once, for the column statistics at set_data
                t = torch.as_tensor(DeviceBuffer(ptr, count), device="cuda")
                torch.cuda.synchronize()
                slots[r] = t.cpu().numpy().copy()
                bar.wait()
                tot = ordered_sum(slots)
                bar.wait()
                t.copy_(torch.from_numpy(tot))
                torch.cuda.synchronize()
                return 0
            e.set_exchange(xfn)
            e.set_data(X[lo:hi], y[lo:hi])
            ptrs[r], _ = e.p2p_mailbox(world)
            bar.wait()
            e.p2p_connect(r, world, dev_ptrs=ptrs)
            e.init_chain(0, beta0)
            bar.wait()
            S, st = e.run(iters, replay_u=replay_u)
            S2, st2 = e.run(3)                    # a second run: the stamps go on, nothing is cleared
            out[r] = (S[0], st, e.state(0), S2[0])
            bar.wait()
            e.close()
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)
            bar.abort()
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    return out


@pytest.mark.parametrize("family,prior", [("gaussian", "normal"), ("binomial", "laplace"), ("poisson", "student_t")])
def test_mailbox_exchange_inside_the_persistent_kernel(family, prior):
    X, y, bt = synth(family, 6002, 4, seed=14)
    m = oracle.make_model(family, **PRIOR_CASES[prior])
    b0 = 0.1 * np.ones(4)
    ref = oracle.run_chain(m, X, y, b0, w=0.3, n_iter=28, seed=9, chain=0)
    for jet in (True, False):
        out = _run_two_shards_mailboxes(family, prior, X, y, b0, 0.3, 25, seed=9, jet=jet)
        for S, st, state, S2 in out:
            assert np.max(np.abs(S - ref["samples"][:25])) <= 1e-9
            assert np.max(np.abs(S2 - ref["samples"][25:])) <= 1e-9
            assert st["launches"] == 1                                   # the whole run is ONE kernel per rank
            assert (st["jet_passes"] > 0) == jet
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][3], out[1][3])     # identical branches on both ranks
        eta_full = np.concatenate([out[0][2][1], out[1][2][1]])
        assert np.max(np.abs(eta_full - X @ out[0][2][0])) < 1e-10


def _p2p_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mcmcglm_b200.multigpu import init_nccl, init_p2p
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    lo, hi = shard_rows(1000, world, rank)
    e = Engine(hi - lo, 3, family="gaussian", w=0.5, n_chains=1, K=6, driver="grid", row_sharded=True, device=rank,
               **PRIOR_CASES["normal"])
    init_nccl(e, rank, world)            # column statistics at set_data
    init_p2p(e, rank, world)             # per-pass exchange: NVLink peer mailboxes
    e.set_data(z["X"][lo:hi], z["y"][lo:hi])
    e.init_chain(0, z["beta0"])
    dist.barrier()
    S, st = e.run(120, replay_u=z["uniforms"])
    q.put((rank, S[0], st["launches"]))
    dist.barrier()
    e.close()
    dist.destroy_process_group()


def test_p2p_mailboxes_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    assert np.array_equal(res[0][1], res[1][1]) and res[0][2] == 1
    assert np.max(np.abs(res[0][1] - z["samples"][1:121])) <= 1e-9


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mcmcglm_b200.multigpu import init_nccl
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    lo, hi = shard_rows(1000, world, rank)
    e = Engine(hi - lo, 3, family="gaussian", w=0.5, n_chains=1, K=6, driver="stepwise", row_sharded=True, device=rank,
               **PRIOR_CASES["normal"])
    init_nccl(e, rank, world)
    e.set_data(z["X"][lo:hi], z["y"][lo:hi])
    e.init_chain(0, z["beta0"])
    S, st = e.run(120, replay_u=z["uniforms"])
    q.put((rank, S[0], st["launches"]))
    e.close()
    dist.destroy_process_group()


def test_nccl_exchange_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(os.path.join(G, "readme_gaussian.npz"))
    assert np.array_equal(res[0][1], res[1][1])
    assert np.max(np.abs(res[0][1] - z["samples"][1:121])) <= 1e-9
