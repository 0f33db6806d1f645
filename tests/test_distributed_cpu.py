"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: row sharding, the
Comm wrapper (sum / product / max / n_global), the common commit schedule for ragged shards, and
the algebra the batched multi-GPU BCA relies on -- replicated state == all-reduce of the shard
states / shard deltas (checked with the CPU oracle's confusion sums)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xcolumns_b200.distributed import Comm, batch_schedule, make_comm, shard_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world, out_dir)
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)


def test_shard_rows_partition():
    for n in (0, 1, 7, 100, 307000):
        for world in (1, 2, 3, 8):
            cuts = [shard_rows(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert batch_schedule(10, 4) == 3 and batch_schedule(0, 4) == 1


def test_single_process_comm_is_a_noop():
    c = make_comm(False)
    t = torch.arange(4, dtype=torch.float64)
    assert c.world == 1 and c.rank == 0 and c.allreduce_sum_(t) is t and c.max_int(5) == 5 and c.n_global(9) == 9


def _comm_basics(rank, world, out_dir):
    c = make_comm(True)
    assert c.world == world and c.rank == rank
    t = torch.full((3, 5), float(rank + 1), dtype=torch.float64)
    c.allreduce_sum_(t)
    assert torch.equal(t, torch.full((3, 5), 3.0, dtype=torch.float64))
    p = torch.full((4,), 0.5 + rank, dtype=torch.float64)
    c.allreduce_prod_(p)
    assert torch.allclose(p, torch.full((4,), 0.75, dtype=torch.float64))
    assert c.max_int(10 + rank) == 11
    assert c.n_global(100 + rank) == 201
    # ragged shards agree on the number of commits, so every rank joins every all-reduce
    n = 1001
    lo, hi = shard_rows(n, rank, world)
    nb = c.max_int(batch_schedule(hi - lo, 64))
    assert nb == batch_schedule(501, 64)
    for _ in range(nb):
        c.allreduce_sum_(torch.zeros(6, dtype=torch.float64))
    assert c.n_allreduce >= nb


def test_comm_basics_gloo(tmp_path):
    _run("_comm_basics", tmp_path)


def _sharded_state(rank, world, out_dir):
    """state(all rows) == all-reduce of state(shard); delta(all changed rows) == all-reduce of the
    shard deltas -- the two identities behind BcaSession.recompute / sweep_batched with comm."""
    from oracle import oracle as orc
    from xcolumns_b200.synth import dense_probs
    c = make_comm(True)
    n, m, k = 600, 90, 4
    eta = dense_probs(n, m, seed=5)
    pred0 = orc.topk_indices_dense(eta, k)[0]
    a = (0.5 + np.random.default_rng(1).random(m)).astype(np.float32)
    pred1 = orc.topk_indices_dense(eta, k, a)[0]     # a different prediction for the same rows

    def sums(rows, pidx):
        p = np.zeros((len(rows), m), dtype=np.float32)
        p[np.arange(len(rows))[:, None], pidx] = 1
        tp, fp, fn, _ = orc.calculate_confusion_matrix(eta[rows], p, skip_tn=True, dtype=np.float64)
        return torch.from_numpy(np.stack([tp, fp, fn]))

    lo, hi = shard_rows(n, rank, world)
    rows = np.arange(lo, hi)
    full0 = sums(np.arange(n), pred0)
    mine0 = sums(rows, pred0[lo:hi])
    state = mine0.clone()
    c.allreduce_sum_(state)
    assert torch.allclose(state, full0, rtol=0, atol=1e-9)
    delta = sums(rows, pred1[lo:hi]) - mine0          # what a batch over my rows would commit
    c.allreduce_sum_(delta)
    assert torch.allclose(state + delta, sums(np.arange(n), pred1), rtol=0, atol=1e-9)
    # n_div must be the global row count on every rank
    assert c.n_global(hi - lo) == n


def test_sharded_state_algebra_gloo(tmp_path):
    _run("_sharded_state", tmp_path)
