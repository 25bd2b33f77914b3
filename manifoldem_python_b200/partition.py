"""Static cost-balanced sharding of projection directions over GPUs.

PDs are independent (the reference hands them to Pool / MPI workers one at a time,
modules/GetDistancesS2.py:110-113, modules/GetDistancesS2_mpi.py:14-15,61-75), so the multi-GPU path is a
partition with NO data-path collective: greedy longest-processing-time onto the least-loaded rank."""


def pd_cost(nS, N):
    """Relative device time of one PD: contraction ~ nS^2 * K (K ~ 1.28 N^2 executed columns, 3 TF32 passes on
    ~820 TF/s) + HBM-bound per-image passes (~15 image-sized round trips at ~5 TB/s)."""
    k = 1.28 * N * N
    return 3.0 * nS * nS * k / 8.2e14 + 15.0 * nS * N * N * 8.0 / 5.0e12


def lpt_partition(costs, n_ranks, speeds=None):
    """Returns a list of n_ranks lists of job indices.  Deterministic: ties broken by index / rank.
    `speeds` (optional, one positive number per rank): relative rate of the ranks — e.g. the host-to-device copy rate each
    GPU reaches while all copy at once (on the round-2 box GPUs 0-3 get 23 GB/s and GPUs 4-7 35 GB/s,
    profiles/r02_h2d_concurrent_8gpu.txt) — a job costs cost / speed on a rank and goes to the rank that finishes it first."""
    if speeds is None:
        speeds = [1.0] * n_ranks
    if len(speeds) != n_ranks or min(speeds) <= 0:
        raise ValueError('speeds: one positive number per rank')
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * n_ranks
    shards = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k] + costs[i] / speeds[k], k))
        shards[r].append(i)
        loads[r] += costs[i] / speeds[r]
    return shards


def counts_by_speed(n_jobs, speeds):
    """Equal jobs over ranks of different speed: how many each rank takes (largest-remainder rounding, sum = n_jobs)."""
    tot = float(sum(speeds))
    raw = [n_jobs * s / tot for s in speeds]
    cnt = [int(x) for x in raw]
    for k in sorted(range(len(speeds)), key=lambda k: (-(raw[k] - cnt[k]), k))[:n_jobs - sum(cnt)]:
        cnt[k] += 1
    return cnt


def round_robin(n_jobs, n_ranks):
    """The reference's MPI split, jobs[j::size] (GetDistancesS2_mpi.py:14-15) — kept for comparison."""
    return [list(range(r, n_jobs, n_ranks)) for r in range(n_ranks)]


def imbalance(costs, shards, speeds=None):
    speeds = speeds or [1.0] * len(shards)
    loads = [sum(costs[i] for i in s) / speeds[r] for r, s in enumerate(shards)]
    mean = sum(loads) / max(1, len(loads))
    return max(loads) / mean if mean > 0 else 1.0
