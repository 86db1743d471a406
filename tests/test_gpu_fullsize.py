"""BASELINE.json configs at their FULL sizes, checked through size-independent properties (the oracle needs
minutes at these sizes): filter == exact predicate, determinism, sharding invariance, count invariants, and
a sampled bit-exact oracle comparison."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from casapose_b200 import synthetic  # noqa: E402
from oracle import ransac_voting_np as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vote(cuda_lib):
    assert torch.cuda.is_available()
    from casapose_b200.pose_estimation import ransac_voting_layer_all_masks

    return ransac_voting_layer_all_masks


def _tile(d, reps):
    return (torch.from_numpy(np.tile(d["mask"], (reps, 1, 1, 1))).cuda(),
            torch.from_numpy(np.tile(d["vertex"], (reps, 1, 1, 1, 1))).cuda())


def _invariants(dbg, hn):
    tn = dbg["tn"].long()
    rounds = dbg["rounds"]
    counts = dbg["counts"]  # [b,oc,mi,hn,vn]
    assert int(counts.min()) >= 0
    assert bool((counts.amax(dim=(2, 3, 4)) <= tn).all()), "a hypothesis got more votes than there are pixels"
    first = counts[:, :, 0]  # round 0
    widx = dbg["win_idx"][:, :, 0].long()  # [b,oc,vn]
    best = first.amax(dim=2)  # [b,oc,vn]
    picked = torch.gather(first, 2, widx.unsqueeze(2)).squeeze(2)
    live = (rounds > 0).unsqueeze(-1)
    assert bool(((picked == best) | ~live).all()), "winner is not an arg-max"
    # arg-max takes the FIRST maximum (:328)
    is_max = first == best.unsqueeze(2)
    first_max = is_max.float().argmax(dim=2)
    assert bool(((first_max == widx) | ~live).all())


def test_config3_13_objects_batch32_full_size(vote):
    """config_13 (13 objects, LM-shaped) voting batch 32 at 480x640, hn = 512."""
    d = synthetic.make_frames(4, 480, 640, synthetic.CONFIG_13_IDS, variant="easy")
    mask, vertex = _tile(d, 8)  # 32 frames; every copy has its own image index -> its own hypothesis stream
    pts, dbg = vote(mask, vertex, 512, seed=21, return_debug=True)
    _invariants(dbg, 512)
    assert int(dbg["stats"][0]) == int((dbg["tn"].long() * dbg["rounds"].long()).sum()) * 9 * 512
    # determinism
    pts2, dbg2 = vote(mask, vertex, 512, seed=21, return_debug=True)
    assert torch.equal(pts, pts2) and torch.equal(dbg["counts"], dbg2["counts"])
    # sharding invariance: frames 8..15 as their own call with image_offset = 8
    part, pdbg = vote(mask[8:16].contiguous(), vertex[8:16].contiguous(), 512, seed=21, image_offset=8, return_debug=True)
    assert torch.equal(part, pts[8:16]) and torch.equal(pdbg["counts"], dbg["counts"][8:16])
    # filter == exact predicate on two full frames
    _, ex = vote(mask[:2].contiguous(), vertex[:2].contiguous(), 512, seed=21, return_debug=True, force_exact=True)
    assert torch.equal(ex["counts"], dbg["counts"][:2]) and torch.equal(ex["win_idx"], dbg["win_idx"][:2])
    # sampled oracle check: frame 5 (= synthetic frame 1 with image index 5), three classes
    i = 5
    for c in (0, 6, 12):
        r = O.ransac_voting_batch(d["mask"][1, :, :, c], d["vertex"][1], 0.99, 0.99, 20, 5, 30000, 512, 9, seed=21, image=i, cls=c)
        assert int(dbg["tn"][i, c]) == r["tn"] and int(dbg["rounds"][i, c]) == r["rounds"]
        for k in range(r["rounds"]):
            assert np.array_equal(dbg["counts"][i, c, k].cpu().numpy(), r["counts"][k])
        assert np.abs(pts[i, c].cpu().numpy() - r["points"]).max() <= 1e-3


def test_config4_batch256_shards_equal_whole(vote):
    """256 LM-O-shaped frames: the result of 8 shards of 32 equals the single call."""
    d = synthetic.make_frames(8, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
    mask, vertex = _tile(d, 32)  # 256 frames, 8.2 GB
    whole = vote(mask, vertex, 512, seed=4)
    assert tuple(whole.shape) == (256, 8, 9, 2) and bool(torch.isfinite(whole).all())
    for s in (0, 3, 7):
        part = vote(mask[s * 32:(s + 1) * 32], vertex[s * 32:(s + 1) * 32], 512, seed=4, image_offset=s * 32)
        assert torch.equal(part, whole[s * 32:(s + 1) * 32])


@pytest.mark.parametrize("hn", [128, 2048])
def test_config5_1080p_cap_active(vote, hn):
    """1080x1920, 8 objects: the 30000-pixel cap (:295-301) is active; hypothesis sweep end points."""
    d = synthetic.make_frames(1, 1080, 1920, synthetic.CONFIG_8_IDS, variant="easy")
    mask, vertex = torch.from_numpy(d["mask"]).cuda(), torch.from_numpy(d["vertex"]).cuda()
    pts, dbg = vote(mask, vertex, hn, seed=8, return_debug=True)
    _invariants(dbg, hn)
    tn0, tn = dbg["tn0"].cpu().numpy()[0], dbg["tn"].cpu().numpy()[0]
    capped = tn0 > 30000
    assert capped.any(), "the synthetic 1080p frame should contain an object above max_num"
    assert (np.abs(tn[capped] - 30000) < 1200).all() and (tn[~capped] == tn0[~capped]).all()
    _, ex = vote(mask, vertex, hn, seed=8, return_debug=True, force_exact=True) if hn == 128 else (None, None)
    if ex is not None:
        assert torch.equal(ex["counts"], dbg["counts"])
    # the capped class against the oracle (down-sampling stream + votes), hn = 128 only (oracle time)
    if hn == 128:
        c = int(np.argmax(tn0))
        r = O.ransac_voting_batch(d["mask"][0, :, :, c], d["vertex"][0], 0.99, 0.99, 20, 5, 30000, hn, 9, seed=8, image=0, cls=c)
        assert r["tn"] == int(tn[c])
        assert np.array_equal(dbg["counts"][0, c, 0].cpu().numpy(), r["counts"][0])
        assert np.abs(pts[0, c].cpu().numpy() - r["points"]).max() <= 1e-3
