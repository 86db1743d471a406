"""Image sharding of the voting path across the GPUs of one node.

The path has no cross-image data (the reference maps independently over images,
/root/reference/casapose/pose_estimation/ransac_voting.py:483), so a batch is partitioned into contiguous
image ranges, one per rank, with NO data-path collective; only the [b, oc, vn, 2] keypoints are gathered.
`image_offset` keeps the Philox hypothesis streams tied to the GLOBAL image index, so the sharded result is
bit-identical to the single-GPU one.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing."""
import torch
import torch.distributed as dist


def shard_bounds(n_images, rank, world):
    """Contiguous range [start, stop) of images owned by `rank`; the first n % world ranks own one more."""
    base, extra = divmod(n_images, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_points(local_points, n_images, group=None):
    """All-gather of the per-rank [b_r, oc, vn, 2] keypoints into [n_images, oc, vn, 2] (image order)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_points
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_images, r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    pad = torch.zeros((width,) + tuple(local_points.shape[1:]), dtype=local_points.dtype, device=local_points.device)
    s, e = sizes[rank]
    pad[: e - s] = local_points
    out = torch.empty((world * width,) + tuple(local_points.shape[1:]), dtype=local_points.dtype, device=local_points.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * width : r * width + (sizes[r][1] - sizes[r][0])] for r in range(world)], dim=0)


def sharded_vote(vote_fn, mask, vertex, round_hyp_num, *, n_images=None, group=None, **kw):
    """Runs `vote_fn` (ransac_voting_layer_all_masks) on this rank's images of a GLOBAL batch and gathers.

    mask / vertex are the rank's own shard (already resident on its GPU, like the network output that
    produced them); n_images is the global batch size (default: world * local)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if n_images is None:
        n_images = world * mask.shape[0]
    start, stop = shard_bounds(n_images, rank, world)
    if stop - start != mask.shape[0]:
        raise ValueError("rank %d owns images [%d,%d) but was given %d" % (rank, start, stop, mask.shape[0]))
    local = vote_fn(mask, vertex, round_hyp_num, image_offset=start, **kw)
    return gather_points(local, n_images, group)
