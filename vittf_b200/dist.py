"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed (NCCL)
for the plumbing.  Pure planning functions (no CUDA needed) + the one collective the path has.

  * ViT slices are independent units: per slicing axis each rank takes a contiguous range of POOLED
    slabs and therefore the slices inside those slabs' AdaptiveAvgPool windows
    ``[floor(o*S/n), ceil((o+1)*S/n))`` -- disjoint whenever S % n == 0, otherwise neighbouring ranks
    both evaluate the shared boundary slice (reads only, no exchange).
  * Each rank writes its slabs into a zero-initialised full-size per-axis buffer; ONE all-reduce(sum)
    per axis assembles the volume.  Supports are disjoint, so every element is x + 0 + ... + 0: exact
    in fp16 in any reduction order.  The z, y, x buffers are then summed in the reference's order.
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


def all_reduce_disjoint(buf, group=None):
    """Sum of per-rank buffers with disjoint supports (exact, see module docstring)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf
