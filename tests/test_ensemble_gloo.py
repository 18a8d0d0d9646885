"""The multi-rank ensemble path on CPU: world_size 2, gloo.  The per-batch runner is backed by the CPU oracle
here (tests may use it as the stand-in executor); what is under test is the partition of the run list over
ranks, the batching and the gather of the reporter rows on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import util as U

T = 6


def _oracle_runner(batch):
    from oracle import oracle_c as O
    from pyproton_oc_b200.engine import REPORTER_NAMES
    out = np.zeros((len(batch), T, len(REPORTER_NAMES)))
    for i, a in enumerate(batch):
        S = O.OracleState(a["initial_state"])
        P = O.make_params(a["source_override"])
        mult = 0.0
        crimes = 0
        for t in range(T):
            mult, totals = S.run_ticks(P, 1 + t, 1, a["seed"], mult)
            crimes += int(totals[0])
            cols = S.cols()
            alive = cols["alive"].astype(bool)
            out[i, t, 0] = crimes
            out[i, t, 7] = int(cols["oc_member"][alive].sum())
            out[i, t, 15] = mult
    return out


def _args():
    from pyproton_oc_b200 import ensemble
    st = U.pre_state(U.load_fixture(os.path.join(U.GOLDEN, "tick_n100_s42_t1.npz")))
    args = ensemble.seed_sweep(5, [{"number_arrests_per_year": 20}, {"number_arrests_per_year": 40}])
    for a in args:
        a["initial_state"] = st
    return args


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from pyproton_oc_b200 import ensemble
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = ensemble.run_ensemble(_args(), _oracle_runner, replicas_per_batch=2, dist=dist)
    if rank == 0:
        q.put(res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_equals_single_process():
    from pyproton_oc_b200 import ensemble
    want = ensemble.run_ensemble(_args(), _oracle_runner, replicas_per_batch=3)
    from pyproton_oc_b200.engine import REPORTER_NAMES
    assert want.shape == (10, T, len(REPORTER_NAMES))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, want)
    # different seeds really gave different runs, same seed + same override gives the same run
    assert not np.array_equal(want[0], want[1])
