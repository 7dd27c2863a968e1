"""Multi-GPU plumbing of the hot path: rays are independent, so every rank takes
a contiguous shard of each batch; parameters are replicated and their gradients summed with ONE
all-reduce per step over the flat fp32 gradient buffer (NCCL on GPUs, gloo in the CPU tests)."""
from typing import Tuple

import torch as th
import torch.distributed as dist


def shard_range(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the rays rank `rank` of `world` takes out of n_rays (contiguous blocks,
    the remainder spread over the first ranks)."""
    base, rem = divmod(n_rays, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(flat_grad: th.Tensor, group=None) -> th.Tensor:
    """In-place sum of the flat gradient buffer over the ranks (one collective per step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def global_mean_scale(world: int) -> float:
    """Every rank's loss is a mean over its own shard; with equal shards the gradient of the
    global mean is the rank sum divided by the world size (barf/model_interpolation.py:508)."""
    return 1.0 / world


def render_rows_sharded(render_fn, n_rows: int, row_width: int, device, group=None, dst: int = 0):
    """Full-image render over the ranks of a process group (rank r takes a contiguous
    block of image rows, the image is assembled on rank `dst` only).  render_fn(row_begin, row_end) ->
    (rows, row_width, 3) fp32 on `device` for this rank's block.  Returns the (n_rows, row_width, 3)
    image on rank `dst` and None elsewhere; one all-gather of equally padded blocks is the only
    collective (rows are independent)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return render_fn(0, n_rows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    begin, end = shard_range(n_rows, rank, world)
    block = (n_rows + world - 1) // world
    mine = th.zeros((block, row_width, 3), device=device, dtype=th.float32)
    if end > begin:
        mine[: end - begin] = render_fn(begin, end)
    gathered = th.empty((world * block, row_width, 3), device=device, dtype=th.float32)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        b, e = shard_range(n_rows, r, world)
        parts.append(gathered[r * block: r * block + (e - b)])
    return th.cat(parts, dim=0)
