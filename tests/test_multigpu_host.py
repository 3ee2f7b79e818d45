"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: dealing chromosomes to ranks and
all-gathering the fitness vectors gives exactly the unsharded result on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deal_is_a_balanced_partition():
    from gaselect_b200 import multigpu, synth
    sizes = [len(s) for s in synth.random_subsets(500, 301, 1, 200, seed=4)]
    for world in (1, 2, 4, 8):
        shares = multigpu.deal(sizes, world)
        allidx = np.sort(np.concatenate(shares))
        assert allidx.tolist() == list(range(len(sizes)))          # a partition
        loads = [sum(sizes[i] ** 2 for i in s) for s in shares]
        assert max(loads) - min(loads) <= max(sizes) ** 2           # LPT bound
        assert all((a == b).all() for a, b in zip(shares, multigpu.deal(sizes, world)))  # deterministic


def test_deal_with_the_kernel_space_cost_and_equal_counts():
    """The kernel-space kernel's cost is flat above 40 columns: the deal balances counts, and with equal_counts every
    rank holds exactly ceil(n / world) chromosomes or one less (equal-sized all-gather buffers)."""
    from gaselect_b200 import multigpu, synth
    sizes = [len(s) for s in synth.random_subsets(1000, 4000, 10, 200, seed=42)]
    cost = multigpu.kspace_cost(sizes)
    for world in (2, 8):
        shares = multigpu.deal(sizes, world, cost=cost, equal_counts=True)
        assert np.sort(np.concatenate(shares)).tolist() == list(range(len(sizes)))
        assert {len(s) for s in shares} == {len(sizes) // world}
        loads = [float(cost[s].sum()) for s in shares]
        assert max(loads) - min(loads) <= 1.5 * float(cost.max())


class _FakeEvaluator:
    """Stands in for the engine in the plumbing test: a deterministic function of the subset."""

    def evaluate_population(self, subsets):
        fit = np.array([-(float(np.sum(np.asarray(s, dtype=np.float64) ** 1.5)) + len(s)) for s in subsets])
        status = np.array([1 if len(s) == 0 else 0 for s in subsets], dtype=np.int32)
        return fit, status


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gaselect_b200 import multigpu, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        subsets = synth.random_subsets(300, 57, 0, 60, seed=8)
        subsets[5] = np.zeros(0, dtype=np.uint32)
        fit, status = multigpu.ShardedEvaluator(_FakeEvaluator()).evaluate_population(subsets)
        want_f, want_s = _FakeEvaluator().evaluate_population(subsets)
        ok = bool((fit == want_f).all() and (status == want_s).all())
        f0, s0 = multigpu.ShardedEvaluator(_FakeEvaluator()).evaluate_population([])
        out.put((rank, ok and len(f0) == 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_evaluation_matches_unsharded_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
