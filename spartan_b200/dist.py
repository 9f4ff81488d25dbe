"""Multi-GPU plumbing for bench.py: one process per GPU, torch.distributed (NCCL on the box, gloo in the CPU tests).
Round 1 shards *units of work* — independent proofs — across ranks: rank r proves the instance with seed r; the only exchange is the
max-over-ranks reduction of the step time and a barrier on both sides of the timed region (no data-path collective)."""
import os


def env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    import torch
    import torch.distributed as dist
    rank, world, local = env()
    if world == 1:
        return rank, world, local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, world, local


def allgather_bytes(b):
    """every rank contributes the same number of bytes; returns the list in rank order (CUDA tensors under NCCL, CPU tensors under gloo)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [bytes(b)]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(b), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


def connect(ctx):
    """make `ctx` (this rank's spartan_b200.Context) one rank of a sharded prover over all ranks of the process group"""
    rank, world, _ = env()
    if world > 1:
        ctx.connect_peers(rank, world, allgather_bytes)
    return ctx


def rank_seed(rank, base_seed=0):
    """instance / tape seed of the proof a rank produces"""
    return base_seed + rank


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values):
    """element-wise max of a list of floats over all ranks"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def aggregate_throughput(units_per_rank, world, max_seconds):
    """whole-job throughput: every rank processed `units_per_rank` units within the slowest rank's time"""
    return units_per_rank * world / max_seconds


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
