"""Population sharding over the GPUs of one node (one process per GPU, torch.distributed).

The reference's only parallelism is MultiThreadedPopulation's contiguous population slices over
host threads (src/MultiThreadedPopulation.cpp:260-262,302-308).  Here the chromosomes of one
generation are dealt to the ranks, every rank scores its share on its own B200 (X, y, the CV
segmentation and the Gram tables are replicated: each engine builds them locally from the same
seed, nothing is exchanged), and the per-generation fitness vectors are all-gathered -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  Chromosomes are independent units: there is no other
data-path collective.
"""
import numpy as np


def deal(sizes, world_size):
    """Longest-processing-time-first deal of chromosomes to ranks.

    Cost proxy of one evaluation: k^2 (the Gram-space SIMPLS matrix-vector products dominate).
    Deterministic (stable sort, ties to the lower rank); returns a list of index arrays, one per
    rank, each ascending."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-(sizes * sizes), kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    shares = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))  # first minimum
        shares[r].append(int(i))
        load[r] += int(sizes[i]) ** 2 + 1
    return [np.asarray(sorted(s), dtype=np.int64) for s in shares]


class ShardedEvaluator:
    """Evaluator front end for N ranks: same ``evaluate_population`` contract as
    gaselect_b200.evaluators.Evaluator, every rank passes the SAME population and receives the
    complete (fitness, status) vectors."""

    def __init__(self, evaluator, group=None, device=None):
        import torch.distributed as dist
        self.ev, self.group, self.device = evaluator, group, device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def evaluate_population(self, subsets):
        import torch
        import torch.distributed as dist
        n = len(subsets)
        shares = deal([len(s) for s in subsets], self.world)
        mine = shares[self.rank]
        fit_l, st_l = self.ev.evaluate_population([subsets[i] for i in mine])
        cap = max(len(s) for s in shares) if n else 0
        dev = self.device if self.device is not None else "cpu"
        send = torch.zeros(2 * cap, dtype=torch.float64, device=dev)
        if len(mine):
            send[: len(mine)] = torch.as_tensor(np.asarray(fit_l, dtype=np.float64), device=dev)
            send[cap: cap + len(mine)] = torch.as_tensor(np.asarray(st_l, dtype=np.float64), device=dev)
        recv = torch.empty(self.world * 2 * cap, dtype=torch.float64, device=dev)
        if cap:
            dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.cpu().numpy().reshape(self.world, 2, cap) if cap else np.zeros((self.world, 2, 0))
        fitness = np.zeros(n, dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        for r, idx in enumerate(shares):
            fitness[idx] = recv[r, 0, : len(idx)]
            status[idx] = recv[r, 1, : len(idx)].astype(np.int32)
        return fitness, status
