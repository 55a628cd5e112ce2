"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed (NCCL)
for the plumbing.  Pure planning functions (no CUDA needed) + the one collective the path has.

  * ViT slices are independent units: per slicing axis each rank takes a contiguous range of POOLED
    slabs and therefore the slices inside those slabs' AdaptiveAvgPool windows
    ``[floor(o*S/n), ceil((o+1)*S/n))`` -- disjoint whenever S % n == 0, otherwise neighbouring ranks
    both evaluate the shared boundary slice (reads only, no exchange).
  * When the pooled slabs divide evenly over the ranks, each rank pools into a COMPACT block (its slabs only) and ONE
    all-gather per axis moves exactly the bytes that are needed; it is launched asynchronously, so the exchange of
    one axis runs under the ViT compute of the next, and a native kernel un-permutes the rank-major result while it
    sums the axes in the reference's order (z, y, x, fp16).  Otherwise (uneven slabs) each rank writes into a
    zero-initialised full-size buffer and ONE all-reduce(sum) per axis assembles the volume: supports are disjoint,
    so every element is x + 0 + ... + 0, exact in fp16 in any reduction order.
  * Similarity / labels shard over z-slabs of the OUTPUT grid; with the feature volume replicated by
    the all-reduce above there is no halo exchange.
"""
import math


def slab_range(n_out, world, rank):
    """Contiguous range of pooled slabs [o0, o1) owned by `rank` (may be empty when world > n_out)."""
    return n_out * rank // world, n_out * (rank + 1) // world


def slices_for_slabs(n_slices, n_out, o0, o1):
    """Slices [a, b) covering the AdaptiveAvgPool windows of slabs [o0, o1)."""
    if o1 <= o0:
        return 0, 0
    return (o0 * n_slices) // n_out, math.ceil(o1 * n_slices / n_out)


def z_range(depth, world, rank):
    """Output z-slab [z0, z1) of `rank` for the similarity / label stage."""
    return depth * rank // world, depth * (rank + 1) // world


def x_range(width, world, rank):
    """Output x-slab [x0, x1) of `rank`: the similarity-only workloads shard along the slowest axis, so that pass 1
    (dots + Gram over the low-res planes under the slab +- 1) shards with the maps."""
    return width * rank // world, width * (rank + 1) // world


def all_reduce_disjoint(buf, group=None):
    """Sum of per-rank buffers with disjoint supports (exact, see module docstring)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def even_slabs(n_out, world):
    """True when every rank owns the same number (> 0) of pooled slabs: the all-gather path applies."""
    return world > 1 and n_out % world == 0


def gather_blocks(block, group=None, async_op=False):
    """All-gather of the ranks' compact per-axis blocks: returns (staging (world, *block.shape), work handle or None)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    staging = torch.empty((world,) + tuple(block.shape), dtype=block.dtype, device=block.device)
    if dist.get_backend(group) == "nccl":
        work = dist.all_gather_into_tensor(staging, block.contiguous(), group=group, async_op=async_op)
    else:                                                   # gloo (CPU tests): no flat-tensor form for every dtype
        work = dist.all_gather([staging[r] for r in range(world)], block.contiguous(), group=group, async_op=async_op)
    return staging, work


def unpermute_gathered(staging, axis):
    """The layout rule of the merge in plain tensor ops (planning reference for the CPU tests and the parity check of
    vittf_accumulate_gathered_f16): staging (world, D, e0, e1, e2) with the slab axis split over ranks -> (D, fX, fY, fZ)."""
    dim = {"x": 1, "y": 2, "z": 3}[axis]                      # slab axis inside one (D, e0, e1, e2) block
    blocks = [staging[r] for r in range(staging.shape[0])]
    import torch
    return torch.cat(blocks, dim=dim)
