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


def gram_cost(sizes):
    """Cost proxy of one evaluation on the per-fit Gram-space kernel: k^2 (the matrix-vector products dominate)."""
    sizes = np.asarray(sizes, dtype=np.float64)
    return sizes * sizes + 1.0


def kspace_cost(sizes, min_k=40):
    """Cost proxy where the kernel-space kernel K1k takes the subsets of more than `min_k` columns (PLS_CV plans with
    n <= 160 rows and <= 10 components): its cost does not depend on the subset size but for the K = Z Z' build;
    below, the per-fit kernel.  Measured at BASELINE config 3 (profiles/README.md, ms per 296 chromosomes):
    0.59 at k = 16, 0.80 at k = 32, 1.25 ... 1.31 for K1k at k = 48 ... 300."""
    sizes = np.asarray(sizes, dtype=np.float64)
    return np.where(sizes > min_k, 1.24 + 0.00023 * sizes, 0.38 + 0.0131 * sizes)


def deal(sizes, world_size, cost=None, equal_counts=False):
    """Longest-processing-time-first deal of chromosomes to ranks.

    `cost`: per-chromosome cost proxy (default gram_cost: k^2).  `equal_counts`: no rank takes more than
    ceil(n / world_size) chromosomes (equal-sized result buffers for the all-gather).
    Deterministic (stable sort, ties to the lower rank); returns a list of index arrays, one per
    rank, each ascending."""
    sizes = np.asarray(sizes, dtype=np.int64)
    c = gram_cost(sizes) if cost is None else np.asarray(cost, dtype=np.float64)
    order = np.argsort(-c, kind="stable")
    cap = -(-len(sizes) // world_size) if equal_counts else len(sizes)
    load = np.zeros(world_size, dtype=np.float64)
    shares = [[] for _ in range(world_size)]
    for i in order:
        open_ranks = [r for r in range(world_size) if len(shares[r]) < cap]
        r = min(open_ranks, key=lambda q: (load[q], q))  # first minimum among the ranks with room
        shares[r].append(int(i))
        load[r] += c[i]
    return [np.asarray(sorted(s), dtype=np.int64) for s in shares]


class ShardedEvaluator:
    """Evaluator front end for N ranks: same ``evaluate_population`` contract as
    gaselect_b200.evaluators.Evaluator, every rank passes the SAME population and receives the
    complete (fitness, status) vectors."""

    def __init__(self, evaluator, group=None, device=None, cost=None):
        """cost: callable sizes -> per-chromosome cost proxy (gram_cost by default; kspace_cost for PLS_CV plans the
        kernel-space kernel takes)."""
        import torch.distributed as dist
        self.ev, self.group, self.device, self.cost = evaluator, group, device, cost
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def evaluate_population(self, subsets):
        import torch
        import torch.distributed as dist
        n = len(subsets)
        sizes = [len(s) for s in subsets]
        shares = deal(sizes, self.world, cost=None if self.cost is None else self.cost(sizes))
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
