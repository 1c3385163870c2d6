"""Multi-GPU plumbing: one process per GPU (torchrun), tiles shard embarrassingly, one collective in fit().

* ``shard_range(B, rank, world)``: contiguous split of the batch dimension -- rank r owns tiles
  ``[r*ceil(B/G), min(B,(r+1)*ceil(B/G)))``.  ``transform`` needs no data-path collective.
* ``share_fit_statistics(vec, src, group)``: the single ``all_reduce(SUM)`` of the fitted target statistics
  (2x3 stain matrix + 1x2 maxC for the extractive normalisers, 3 means + 3 stds for Reinhard): ``src`` contributes the
  values, every other rank contributes zeros.  NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch


def shard_range(B, rank, world):
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def share_fit_statistics(vec, src=0, group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return vec
    if dist.get_rank(group) != src:
        vec = torch.zeros_like(vec)
    backend = dist.get_backend(group)
    buf = vec.cuda() if backend == "nccl" else vec.cpu().clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf.to(vec.device)
