"""One process per GPU, torch.distributed for the plumbing (mirror of MuseDiffusion/utils/dist_util.py:58-152).
The sampling loop itself runs no collective: sequences are independent, so ranks take disjoint shards; NCCL is used
only to broadcast weights once and to gather decoded token ids."""
import os

import torch
import torch.distributed as dist


def setup(backend=None):
    """dist_util.setup_dist (:58-86).  Returns (rank, world_size, device); single-process when RANK is unset."""
    if "RANK" in os.environ and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group(backend=backend, init_method="env://")
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return rank, world, dev


def barrier():
    if dist.is_initialized():
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def shutdown():
    """Tear the process group down (no-op when single-process)."""
    if dist.is_initialized():
        dist.destroy_process_group()


def shard_range(n, rank, world):
    """contiguous shard [lo, hi) of n sequences for this rank."""
    per, extra = divmod(n, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def broadcast_model(model, src=0):
    """dist_util.sync_params (:141-152): one broadcast per parameter/buffer from `src`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for p in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(p.data, src)
    # writes through `.data` do not bump the tensors' version counters, which key the packed bf16 weights and the split
    # embedding: rebuild them explicitly so that no rank keeps serving its pre-broadcast weights
    from . import ops
    ops._split_cache.clear()
    if hasattr(model, "weight_pack") and next(model.parameters()).is_cuda:
        model.weight_pack(force=True)


def all_gather_tokens(tokens):
    """[B_local, L] integer token ids of every rank -> [sum B_local, L] on every rank (equal B_local per rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tokens
    out = [torch.empty_like(tokens) for _ in range(dist.get_world_size())]
    dist.all_gather(out, tokens.contiguous())
    return torch.cat(out, dim=0)


def gather_objects(obj):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
