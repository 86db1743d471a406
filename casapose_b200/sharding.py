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


class PendingGather:
    """An all-gather in flight: `wait()` returns the [n_images, oc, vn, 2] keypoints in image order."""

    def __init__(self, work, out, sizes, width, local):
        self._work, self._out, self._sizes, self._width, self._local = work, out, sizes, width, local

    def wait(self):
        if self._work is None:
            return self._local
        self._work.wait()
        w = self._width
        return torch.cat([self._out[r * w : r * w + (e - s)] for r, (s, e) in enumerate(self._sizes)], dim=0)


def gather_points_async(local_points, n_images, group=None, stream=None, after=None):
    """Non-blocking variant of gather_points: lets the gather of step i overlap the voting of step i + 1
    (the exchange is 576 bytes per frame, so its cost is pure latency).

    stream / after: run the collective on a side CUDA stream that waits only for the event `after` (recorded
    right behind the step that produced `local_points`), so it neither waits for nor delays later work that is
    already queued on the compute stream."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return PendingGather(None, None, None, 0, local_points)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_images, r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    s, e = sizes[rank]

    def issue():
        if e - s == width:
            pad = local_points.contiguous()
        else:
            pad = torch.zeros((width,) + tuple(local_points.shape[1:]), dtype=local_points.dtype, device=local_points.device)
            pad[: e - s] = local_points
        out = torch.empty((world * width,) + tuple(local_points.shape[1:]), dtype=local_points.dtype, device=local_points.device)
        return dist.all_gather_into_tensor(out, pad, group=group, async_op=True), out

    if stream is not None and local_points.is_cuda:
        with torch.cuda.stream(stream):
            if after is not None:
                stream.wait_event(after)
            local_points.record_stream(stream)
            work, out = issue()
            work.wait()  # orders the side stream (not the host, not the compute stream) behind the collective
        return _StreamGather(stream, out, sizes, width)
    work, out = issue()
    return PendingGather(work, out, sizes, width, local_points)


class _StreamGather:
    def __init__(self, stream, out, sizes, width):
        self._stream, self._out, self._sizes, self._width = stream, out, sizes, width

    def wait(self):
        self._stream.synchronize()
        w = self._width
        return torch.cat([self._out[r * w : r * w + (e - s)] for r, (s, e) in enumerate(self._sizes)], dim=0)


class AbiGather:
    """The per-step keypoint all-gather through the library's C ABI (casa_comm_init / casa_allgather_points_overlapped):
    NCCL is driven by the library, torch.distributed only carries the 128-byte NCCL id to the other ranks once
    (any other channel would do: the reference's TensorFlow replicas have no torch process group).

    The exchange runs on the handle's own gather stream behind the vote that produced the rows, so the gather of
    step i overlaps the voting of step i + 1 and no compute stream ever waits for a peer.  `slots` result buffers
    alternate; the vote of step i writes straight into `buffer(i)` (its `out=` argument)."""

    def __init__(self, shape, device, world, rank, group=None, slots=2):
        import ctypes as C

        from . import _lib

        self.world, self.rank, self.device = world, rank, device
        self.lib = _lib.lib()
        self.hdl = _lib.handle(device.index, torch.cuda.current_stream(device).cuda_stream)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(self.lib.casa_nccl_unique_id(buf))
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if world > 1:
            carrier = ident.to(device) if dist.get_backend(group) == "nccl" else ident
            dist.broadcast(carrier, src=0, group=group)
            ident = carrier.cpu()
        raw = (C.c_char * 128).from_buffer_copy(ident.numpy().tobytes())
        _lib.check(self.lib.casa_comm_init(self.hdl, raw, rank, world))
        self.inp = [torch.zeros(shape, dtype=torch.float32, device=device) for _ in range(slots)]
        self.out = [torch.zeros((world * shape[0],) + tuple(shape[1:]), dtype=torch.float32, device=device) for _ in range(slots)]
        self.used = [False] * slots
        self._check = _lib.check

    def buffer(self, step):
        """Result tensor the vote of `step` must write; the compute stream first waits (on the GPU) for the gather
        that last read it — two steps ago, long finished."""
        k = step % len(self.inp)
        if self.used[k]:
            self._check(self.lib.casa_gather_wait(self.hdl, k, torch.cuda.current_stream(self.device).cuda_stream))
        return self.inp[k]

    def launch(self, step):
        """Queue the gather of `step` behind everything queued on the current stream (its vote)."""
        k = step % len(self.inp)
        self._check(self.lib.casa_allgather_points_overlapped(
            self.hdl, self.inp[k].data_ptr(), self.out[k].data_ptr(), self.inp[k].numel(),
            torch.cuda.current_stream(self.device).cuda_stream, k))
        self.used[k] = True
        return _AbiGatherResult(self, k)

    def barrier(self):
        """Device-side barrier on the compute stream: an all-gather of the first result buffer."""
        self._check(self.lib.casa_allgather_points(self.hdl, None, self.inp[0].data_ptr(), self.out[0].data_ptr(),
                                                   self.inp[0].numel(), torch.cuda.current_stream(self.device).cuda_stream))

    def close(self):
        self._check(self.lib.casa_comm_destroy(self.hdl))


class _AbiGatherResult:
    def __init__(self, owner, k):
        self._owner, self._k = owner, k

    def wait(self):
        self._owner._check(self._owner.lib.casa_gather_wait(self._owner.hdl, self._k, -1))
        return self._owner.out[self._k]


def sharded_vote(vote_fn, mask, vertex, round_hyp_num, *, n_images=None, group=None, async_gather=False, **kw):
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
    if async_gather:
        return gather_points_async(local, n_images, group)
    return gather_points(local, n_images, group)
