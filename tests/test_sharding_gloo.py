"""world_size-2 gloo test (CPU) of the N > 1 host logic: contiguous image shards + gather give exactly the
unsharded result.  The compute stand-in on CPU is the oracle (tests may use it); on GPUs bench.py passes the
CUDA path through the same `sharded_vote`."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casapose_b200 import sharding, synthetic


def test_shard_bounds_cover_the_batch_once():
    for n in (1, 5, 16, 17, 256):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_vote(mask, vertex, hn, image_offset=0, **kw):
    from oracle import ransac_voting_np as O

    return torch.from_numpy(O.ransac_voting_layer_all_masks(mask.numpy(), vertex.numpy(), hn, image_offset=image_offset, **kw))


def _worker(rank, world, port, n_images, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synthetic.make_frames(n_images, 60, 80, (1, 5), variant="easy")
    s, e = sharding.shard_bounds(n_images, rank, world)
    pts = sharding.sharded_vote(_oracle_vote, torch.from_numpy(d["mask"][s:e]), torch.from_numpy(d["vertex"][s:e]), 32,
                                n_images=n_images, seed=3)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), pts.numpy())
    pend = sharding.sharded_vote(_oracle_vote, torch.from_numpy(d["mask"][s:e]), torch.from_numpy(d["vertex"][s:e]), 32,
                                 n_images=n_images, seed=3, async_gather=True)
    assert torch.equal(pend.wait(), pts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_the_unsharded_batch(tmp_path):
    n_images = 3  # uneven split: rank 0 owns 2 images, rank 1 owns 1
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    d = synthetic.make_frames(n_images, 60, 80, (1, 5), variant="easy")
    full = _oracle_vote(torch.from_numpy(d["mask"]), torch.from_numpy(d["vertex"]), 32, seed=3).numpy()
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == full.shape and np.array_equal(got, full)
