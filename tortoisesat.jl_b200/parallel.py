"""Multi-GPU plumbing: one process per GPU, trials sharded with NO data-path collective
(trials are independent: the loop body of src/monte_carlo.jl:118-235 touches only index i).
The only exchange is at the end of a run: an all-gather of the 64-byte per-trial outcome
records and an all-reduce(sum) of the statistics vector (NCCL over NVLink on GPUs; gloo in
the CPU tests)."""
import numpy as np

from .host import OUTCOME_DTYPE

STAT_FIELDS = ("n_trials", "n_converged", "n_no_cutoff", "n_fail_slew", "sum_slew_time", "sum_slew_time_sq", "sum_t_final",
               "sum_inner_iters", "sum_ls_rollouts", "sum_knots", "flops")


def shard_range(n_total, rank, world):
    """Contiguous block of trials owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def interleaved_shard(n_total, rank, world):
    """Trial t -> rank t mod world (ragged horizons balance better when interleaved)."""
    return np.arange(rank, int(n_total), int(world), dtype=np.int64)


def stats_vector(st):
    return np.array([float(getattr(st, k)) for k in STAT_FIELDS], dtype=np.float64)


def gather_outcomes(out_local, device=None, group=None):
    """all-gather of outcome records; every rank must pass the same number of records.
    Returns an array of world*len(out_local) records ordered by rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out_local.copy()
    world = dist.get_world_size(group)
    raw = np.frombuffer(np.ascontiguousarray(out_local).tobytes(), dtype=np.uint8).copy()
    loc = torch.from_numpy(raw)
    if device is not None:
        loc = loc.to(device)
    full = torch.empty(world * loc.numel(), dtype=torch.uint8, device=loc.device)
    dist.all_gather_into_tensor(full, loc, group=group)
    return np.frombuffer(full.cpu().numpy().tobytes(), dtype=OUTCOME_DTYPE).copy()


def reduce_stats(vec, device=None, group=None):
    """all-reduce(sum) of the statistics vector."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(vec, dtype=np.float64).copy()
    t = torch.tensor(np.asarray(vec, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
