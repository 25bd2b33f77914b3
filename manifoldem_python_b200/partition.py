"""Static cost-balanced sharding of projection directions over GPUs.

PDs are independent (the reference hands them to Pool / MPI workers one at a time,
modules/GetDistancesS2.py:110-113, modules/GetDistancesS2_mpi.py:14-15,61-75), so the multi-GPU path is a
partition with NO data-path collective: greedy longest-processing-time onto the least-loaded rank."""
import heapq


def pd_cost(nS, N):
    """Relative device time of one PD: contraction ~ nS^2 * K (K ~ 1.28 N^2 executed columns, 3 TF32 passes on
    ~820 TF/s) + HBM-bound per-image passes (~15 image-sized round trips at ~5 TB/s)."""
    k = 1.28 * N * N
    return 3.0 * nS * nS * k / 8.2e14 + 15.0 * nS * N * N * 8.0 / 5.0e12


def lpt_partition(costs, n_ranks):
    """Returns a list of n_ranks lists of job indices.  Deterministic: ties broken by index / rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    shards = [[] for _ in range(n_ranks)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return shards


def round_robin(n_jobs, n_ranks):
    """The reference's MPI split, jobs[j::size] (GetDistancesS2_mpi.py:14-15) — kept for comparison."""
    return [list(range(r, n_jobs, n_ranks)) for r in range(n_ranks)]


def imbalance(costs, shards):
    loads = [sum(costs[i] for i in s) for s in shards]
    mean = sum(loads) / max(1, len(loads))
    return max(loads) / mean if mean > 0 else 1.0
