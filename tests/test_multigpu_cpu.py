"""Host-side logic of the multi-GPU modes on CPU: row sharding, and the rank-ordered exchange over a real
world_size-2 gloo process group (bit-identical totals on every rank)."""
import os
import socket
import numpy as np
import pytest
import oracle
from helpers import synth, PRIOR_CASES
from mcmcglm_b200.multigpu import shard_rows, ordered_sum


def test_shard_rows_partition():
    for n in (1, 2, 7, 1000, 1001, 50_000_000):
        for world in (1, 2, 3, 8):
            blocks = [shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(lo % 2 == 0 for lo, _ in blocks)
            sizes = [hi - lo for lo, hi in blocks]
            if n >= 2 * world:
                assert max(sizes) - min(sizes) <= 2


def test_sharded_log_likelihood_sums_to_the_whole():
    # what the row-sharded mode relies on: f = sum over shards of the shard log-likelihoods + prior
    X, y, bt = synth("poisson", 5001, 4, seed=6)
    m = oracle.make_model("poisson", **PRIOR_CASES["student_t"])
    eta = oracle.init_eta(X, bt)
    full = oracle.log_potential(m, X, y, bt, eta, 2, [bt[2] + 0.05])[0]
    prior = oracle.log_prior_density(m, np.r_[bt[:2], bt[2] + 0.05, bt[3:]])
    parts = []
    for r in range(3):
        lo, hi = shard_rows(5001, 3, r)
        f = oracle.log_potential(m, X[lo:hi], y[lo:hi], bt, eta[lo:hi], 2, [bt[2] + 0.05])[0]
        parts.append([f - prior])
    assert abs(ordered_sum(parts)[0] + prior - full) <= 1e-12 * abs(full)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mcmcglm_b200.multigpu import ordered_sum_exchange
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    local = rng.standard_normal(64) * 10.0 ** rng.integers(-8, 8, 64)     # badly scaled: order matters in fp64
    local[3] = -np.inf if rank == 1 else local[3]
    t = torch.from_numpy(local.copy())
    ordered_sum_exchange(t)
    q.put((rank, local, t.numpy().copy()))
    dist.destroy_process_group()


def test_gloo_world2_exchange_is_bit_identical_and_rank_ordered():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = ordered_sum([res[0][1], res[1][1]])
    assert np.array_equal(res[0][2], res[1][2], equal_nan=True)          # same bits on both ranks
    assert np.array_equal(res[0][2], expect, equal_nan=True)             # = the rank-ordered sum
    assert res[0][2][3] == -np.inf                                       # -Inf propagates, never NaN
