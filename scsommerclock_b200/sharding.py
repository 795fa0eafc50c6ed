"""How the PT test spreads over the GPUs of one node (one process per GPU).

Two shapes of work (BASELINE north_star):
  * one large dataset: SNV rows are block-partitioned across ranks, each rank
    places its rows against the whole tree, and the per-branch soft weights (a
    vector of 2n doubles) are summed with ONE all-reduce -- the only exchange the
    path has;
  * replicate sweeps (many independent VCF/tree pairs): whole replicates are
    dealt to ranks, no communication at all.
torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests).
"""
import os


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of ``n`` items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replicate_shard(n_replicates, rank, world):
    """Indices of the replicates rank ``rank`` runs (round-robin keeps uneven costs balanced)."""
    return list(range(rank, int(n_replicates), int(world)))


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    """Joins the torchrun rendezvous if there is one.  Returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                # SCS_DIST_BACKEND=gloo lets several ranks share one GPU (NCCL refuses duplicate devices): used by the tests
                backend = os.environ.get("SCS_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(local)
                dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_soft_weights(y):
    """In-place sum of the branch-weight vector over ranks (a no-op on one rank)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if y.is_cuda and dist.get_backend() == "gloo":
            h = y.cpu()                      # gloo: stage through the host
            dist.all_reduce(h, op=dist.ReduceOp.SUM)
            y.copy_(h)
        else:
            dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], dtype=torch.float64, device=None if dist.get_backend() == "gloo" else device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
