"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCE over the numpy TensorFlow stand-in.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only (it reads
/root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden ransac_easy ls_plain   # some

The reference modules are imported unmodified from /root/reference; ``tensorflow`` and
``tensorflow_addons`` resolve to oracle/tf_standin/ (TensorFlow 2.9.1 is not installable here).
Random numbers: ``tf.random.uniform`` (ransac_voting.py:296, :319) is answered with the explicit Philox
streams of oracle/philox_np.py keyed by (seed, image, class, round), where image / class are the
iteration indices of the reference's two nested ``tf.map_fn`` calls (:483, :442).  Vote counts are the
operand of the reference's own ``tf.argmax`` call (:328), recorded by the stand-in.

Every .npz holds the inputs (or, for the full-size cases, the synthetic-generator arguments plus a
SHA-256 of the generated inputs), the parameters, and what the reference code returned.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")
F = np.float32
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.golden_inputs import (degenerate_scene, ls_filter_inputs, ls_inputs, metric_scene, pose_inputs,  # noqa: E402
                                  pvnet_inputs, ransac_inputs, sha, unmap_inputs)


def load_reference():
    """Import the reference's modules with the stand-in as ``tensorflow``."""
    standin = os.path.join(ROOT, "oracle", "tf_standin")
    for p in (REFERENCE, standin):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tensorflow as tf  # the stand-in

    assert tf.__version__.endswith("numpy-standin"), "a real TensorFlow is importable: use it instead"
    mods = {name: importlib.import_module("casapose.pose_estimation." + name)
            for name in ("ransac_voting", "voting_layers_2d", "pose_evaluation", "bpnp_layers")}
    for m in mods.values():
        assert m.__file__.startswith(REFERENCE), m.__file__
    return tf, mods


def load_reference_ls_torch():
    """voting_layers_2d.py of the reference, loaded a second time with oracle/tf_standin_torch as ``tensorflow``
    (torch.autograd then differentiates the graph the reference's own code builds)."""
    import importlib.util

    standin = os.path.join(ROOT, "oracle", "tf_standin_torch")
    saved = {k: sys.modules.pop(k, None) for k in ("tensorflow", "tensorflow_addons")}
    sys.path.insert(0, standin)
    try:
        import tensorflow as tft

        assert tft.__version__.endswith("torch-standin")
        path = os.path.join(REFERENCE, "casapose", "pose_estimation", "voting_layers_2d.py")
        spec = importlib.util.spec_from_file_location("reference_voting_layers_2d_torch", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(standin)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return mod


class PhiloxProvider:
    """Answers tf.random.uniform with oracle/philox_np.py streams; counts rounds per (image, class)."""

    def __init__(self, seed, vn, image_offset=0):
        self.seed, self.vn, self.image_offset = seed, vn, image_offset
        self.rounds = {}

    def __call__(self, shape, minval, maxval, dtype, map_index):
        from oracle import philox_np

        image, cls = map_index
        if np.dtype(dtype) == np.float32:  # selection [h,w]  (:296)
            h, w = shape
            return philox_np.draw_selection(self.seed, image + self.image_offset, cls, h, w)
        hn, vn, two = shape  # idxs [hn,vn,2]  (:319)
        assert two == 2 and vn == self.vn and int(minval) == 0
        r = self.rounds.get(map_index, 0)
        self.rounds[map_index] = r + 1
        return philox_np.draw_idxs(self.seed, image + self.image_offset, cls, r, hn, vn, int(maxval))


def run_reference_ransac(tf, mods, mask, vertex, hn, seed, **kw):
    """ransac_voting_layer_all_masks of the reference -> points, rounds [b,oc], counts {(i,c,r): [hn,vn]}."""
    b, _, _, oc = mask.shape
    vn = vertex.shape[3]
    prov = PhiloxProvider(seed, vn)
    tf.random.provider = prov
    del tf.taps[:]
    pts = mods["ransac_voting"].ransac_voting_layer_all_masks(mask, vertex, hn, **kw)
    rounds = np.zeros((b, oc), np.int32)
    counts = {}
    for name, idx, operand in tf.taps:
        if name == "argmax" and len(idx) == 2:
            i, c = idx
            counts["counts_%d_%d_%d" % (i, c, rounds[i, c])] = operand.astype(np.int32)
            rounds[i, c] += 1
    for (i, c), r in prov.rounds.items():
        assert rounds[i, c] == r
    return np.asarray(pts, F), rounds, counts


def labels_of(mask):
    lab = np.zeros(mask.shape[:3], np.uint8)
    for c in range(mask.shape[3]):
        assert not (lab[mask[..., c] != 0]).any(), "one-hot masks only"
        lab[mask[..., c] != 0] = c + 1
    return lab


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.0), flush=True)


# ------------------------------------------------------------------------------------------------ cases
def _ransac_case(tf, mods, name, mask, vertex, hn, seed, gen=None, **kw):
    """gen = None: the inputs are stored; otherwise gen = ransac_inputs(**gen) regenerates them."""
    digest = sha(mask, vertex)
    pts, rounds, counts = run_reference_ransac(tf, mods, mask, vertex, hn, seed, **kw)
    assert sha(mask, vertex) == digest, "the reference code modified its inputs"
    tf.ACCUMULATE = "native"  # numpy's float32 summation order in the refinement (:361-362): sensitivity record
    pts_native, rounds_native, _ = run_reference_ransac(tf, mods, mask, vertex, hn, seed, **kw)
    tf.ACCUMULATE = "float64"
    assert np.array_equal(rounds, rounds_native)
    out = dict(points=pts, rounds=rounds, hn=np.int32(hn), seed=np.int64(seed), input_sha=np.array(digest),
               points_f32_order=pts_native,
               params=np.array(repr(dict(kw))), **counts)
    if gen is None:
        out.update(labels=labels_of(mask), vertex=vertex)
    else:
        out["gen"] = np.array(repr(gen))
    save(name, **out)


def case_ransac_easy(tf, mods):
    _ransac_case(tf, mods, "ransac_easy", *ransac_inputs(2, 96, 128, (1, 5, 6)), 64, 7)


def case_ransac_hard(tf, mods):
    """sigma 20 deg, 60 % random directions: several rounds (stop test :344-346, strict-< update :336)."""
    _ransac_case(tf, mods, "ransac_hard", *ransac_inputs(2, 96, 128, (1, 5, 6), variant="hard"), 64, 7, max_iter=6)


def case_ransac_cap(tf, mods):
    """max_num below the class sizes: the selection stream and the float32 ratio of :295-301."""
    _ransac_case(tf, mods, "ransac_cap", *ransac_inputs(1, 120, 160, (1, 5, 6, 8)), 64, 5, max_num=150)


def case_ransac_params(tf, mods):
    """Non-default parameters: inlier_thresh 0.97, confidence 0.9999 (more rounds), min_num 100 (gates small classes)."""
    _ransac_case(tf, mods, "ransac_params", *ransac_inputs(2, 96, 128, (1, 5, 6), variant="hard", seed=99), 32, 21,
                 inlier_thresh=0.97, confidence=0.9999, max_iter=8, min_num=100)


def case_ransac_degenerate(tf, mods):
    _ransac_case(tf, mods, "ransac_degenerate", *degenerate_scene(), 32, 3, max_iter=3)


def case_ransac_full(tf, mods):
    """BASELINE config 2 shape, two frames: 480x640, 8 objects, 512 hypotheses."""
    from casapose_b200 import synthetic

    gen = dict(b=2, h=480, w=640, ids=synthetic.CONFIG_8_IDS)
    _ransac_case(tf, mods, "ransac_full_480x640", *ransac_inputs(**gen), 512, 1237, gen=gen)


def case_ransac_13obj_hard(tf, mods):
    """BASELINE config 3 shape (13 LM objects), hard variant: multi-round at full resolution (128 hypotheses per round)."""
    from casapose_b200 import synthetic

    gen = dict(b=1, h=480, w=640, ids=synthetic.CONFIG_13_IDS, variant="hard")
    _ransac_case(tf, mods, "ransac_13obj_hard_480x640", *ransac_inputs(**gen), 128, 1237, gen=gen, max_iter=4)


def case_ransac_1080p(tf, mods):
    """BASELINE config 5 shape: 1080x1920, 8 objects, 128 hypotheses; objects above 30000 px take the cap (:295)."""
    from casapose_b200 import synthetic

    gen = dict(b=1, h=1080, w=1920, ids=synthetic.CONFIG_8_IDS)
    mask, vertex = ransac_inputs(**gen)
    assert (mask.sum((1, 2)) > 30000).any(), "no class above max_num: the cap is not exercised"
    _ransac_case(tf, mods, "ransac_1080p_cap", mask, vertex, 128, 1237, gen=gen)


def _run_ls(mods, seg, direct, conf, **layer_kw):
    L = mods["voting_layers_2d"].CoordLSVotingWeighted
    layer = L("ls", seg.shape[-1], num_points=9, **layer_kw)
    return np.asarray(layer([seg, direct, conf]), F)


def case_ls_plain(tf, mods):
    seg, direct, conf = ls_inputs(2, 64, 80, (1, 5, 6))
    save("ls_plain", seg=seg, direct=direct, conf=conf, input_sha=np.array(sha(seg, direct, conf)),
         points=_run_ls(mods, seg, direct, conf),
         points_sigmoid=_run_ls(mods, seg, direct, conf, sigmoid_weights=True))


def case_ls_filter(tf, mods):
    """filter_estimates: connected components, bincount < 50 -> 0, top_k, label 0 competing (:43-79)."""
    seg, direct, conf = ls_filter_inputs()
    save("ls_filter", seg=seg, direct=direct, conf=conf, input_sha=np.array(sha(seg, direct, conf)),
         points=_run_ls(mods, seg, direct, conf, filter_estimates=True),
         points_second=_run_ls(mods, seg, direct, conf, filter_estimates=True, output_second_largest_component=True))


def case_ls_grad(tf, mods):
    """Gradient of L = sum(layer(seg, direct, w) * g) w.r.t. direct and w from the reference's own forward code
    (torch.autograd over oracle/tf_standin_torch), plain softplus weights and filter_estimates + sigmoid weights."""
    import torch

    from oracle.golden_inputs import ls_grad_inputs

    seg, direct, conf, g = ls_grad_inputs()
    mod = load_reference_ls_torch()
    out = {}
    for tag, kw in (("plain", {}), ("filter_sigmoid", dict(filter_estimates=True, sigmoid_weights=True))):
        layer = mod.CoordLSVotingWeighted("ls", seg.shape[-1], num_points=9, **kw)
        td = torch.tensor(direct, requires_grad=True)
        tw = torch.tensor(conf, requires_grad=True)
        res = layer([torch.tensor(seg), td, tw])
        (res * torch.tensor(g)).sum().backward()
        ref = _run_ls(mods, seg, direct, conf, **kw)  # the numpy stand-in's forward of the same code
        assert np.abs(res.detach().numpy() - ref).max() < 1e-3
        out["points_" + tag] = res.detach().numpy().astype(F)
        out["grad_direct_" + tag] = td.grad.numpy().astype(F)
        out["grad_conf_" + tag] = tw.grad.numpy().astype(F)
        assert np.isfinite(out["grad_direct_" + tag]).all() and np.isfinite(out["grad_conf_" + tag]).all()
    save("ls_grad", input_sha=np.array(sha(seg, direct, conf, g)), **out)


def case_ls_full(tf, mods):
    """BASELINE config 1 shape: one 480x640 frame, 8 objects, with and without filter_estimates."""
    from casapose_b200 import synthetic

    gen = dict(b=1, h=480, w=640, ids=synthetic.CONFIG_8_IDS)
    seg, direct, conf = ls_inputs(**gen)
    save("ls_full_480x640", gen=np.array(repr(gen)), input_sha=np.array(sha(seg, direct, conf)),
         points=_run_ls(mods, seg, direct, conf), points_filter=_run_ls(mods, seg, direct, conf, filter_estimates=True))


STAT_NAMES = ["valid_2d", "valid_3d", "valid_pose_count", "false_positive_mask", "err_2d", "err_3d", "missing_object",
              "false_positive_pose"]


def case_pose_eval(tf, mods):
    """estimate_and_evaluate_poses end to end (pose_evaluation.py:11-101): argmax / one-hot pre-step, voting with
    512 hypotheses, offsets un-mapping, OpenCV PnP, ADD and 2-D verdicts."""
    gen = dict(b=2, h=240, w=320, ids=(1, 5, 6, 8, 9, 10, 11, 12), crop=(7.0, 3.0))
    d, cams, offsets, kp3, target_seg, poses_gt, diam = pose_inputs(**gen)
    seed = 11
    tf.random.provider = PhiloxProvider(seed, 9)
    vertex18 = d["vertex"].reshape(gen["b"], gen["h"], gen["w"], 18)
    stats, poses, pts = mods["pose_evaluation"].estimate_and_evaluate_poses(
        d["seg_logits"], target_seg, vertex18, poses_gt, kp3, cams, diam, offsets, min_num=20)
    save("pose_eval", gen=np.array(repr(gen)), input_sha=np.array(sha(d["seg_logits"], vertex18, target_seg, offsets)),
         seed=np.int64(seed), poses=np.asarray(poses, F), points=np.asarray(pts, F),
         **{n: np.asarray(s, F) for n, s in zip(STAT_NAMES, stats)})


def case_pose_eval_pvnet(tf, mods):
    """estimate_and_evaluate_poses with one vector field per class (pose_evaluation.py:38-45) and pose_estimation
    (:222-269) on the same inputs."""
    d, fields, cams, offsets, kp3, target_seg, poses_gt, diam = pvnet_inputs()
    seed = 13
    tf.random.provider = PhiloxProvider(seed, 9)
    stats, poses, pts = mods["pose_evaluation"].estimate_and_evaluate_poses(
        d["seg_logits"], target_seg, fields, poses_gt, kp3, cams, diam, offsets, min_num=20)
    tf.random.provider = PhiloxProvider(seed, 9)
    poses2 = mods["pose_evaluation"].pose_estimation(d["seg_logits"], target_seg, fields, poses_gt, kp3, cams, offsets,
                                                     min_num=20)
    save("pose_eval_pvnet", input_sha=np.array(sha(d["seg_logits"], fields, target_seg, offsets)), seed=np.int64(seed),
         poses=np.asarray(poses, F), points=np.asarray(pts, F), poses_pose_estimation=np.asarray(poses2, F),
         **{n: np.asarray(s, F) for n, s in zip(STAT_NAMES, stats)})


def case_poses_pnp(tf, mods):
    """LS layer -> poses_pnp (pose_evaluation.py:164-217): the (y,x)->(x,y) flip, availability from the hard
    softmax, BPNP_fast forward (OpenCV), Rodrigues, t_z flip."""
    gen = dict(b=1, h=240, w=320, ids=(1, 5, 6, 8, 9, 10, 11, 12), variant="clean")
    d, cams, offsets, kp3, target_seg, poses_gt, diam = pose_inputs(**gen)
    vertex18 = d["vertex"].reshape(gen["b"], gen["h"], gen["w"], 18)
    coords = _run_ls(mods, d["seg_logits"], vertex18, d["conf_logits"], filter_estimates=True)
    poses = np.asarray(mods["pose_evaluation"].poses_pnp(coords, d["seg_logits"], kp3, cams, len(gen["ids"]), min_num=20), F)
    save("poses_pnp", gen=np.array(repr(gen)), input_sha=np.array(sha(d["seg_logits"], vertex18, d["conf_logits"])),
         coords=coords, poses=poses)


def case_unmap(tf, mods):
    """map_offsets (ransac_voting.py:487-504) row by row on random crop / rotation / scale parameters."""
    pts, off = unmap_inputs()
    rv = mods["ransac_voting"]
    out = np.stack([np.asarray(rv.map_offsets(pts[i], 1, off[i]), F) for i in range(len(pts))])
    save("unmap_offsets", points=pts, offsets=off, unmapped=out)


def case_pose_metric(tf, mods):
    """evaluate_poses (ransac_voting.py:628-687) incl. ADD-S on 7862- and 3417-point clouds (:596-621)."""
    s = metric_scene()
    b, oc = s["valid"].shape
    ev = np.tile(s["evaluation_points"][None, :, None], [b, 1, 1, 1, 1])
    cnt = np.tile(s["counts"][None], [b, 1, 1])
    res = mods["ransac_voting"].evaluate_poses(s["poses"], s["poses_gt"], s["points_estimated"], ev, cnt, s["cams"],
                                               s["diameters"], s["valid"], 5.0)
    names = ["err_2d", "err_3d", "valid_2d", "valid_3d", "missing_object", "valid_points_count", "false_positive_pose"]
    save("pose_metric", input_sha=np.array(sha(*[s[k] for k in sorted(s)])), **{n: np.asarray(r, F) for n, r in zip(names, res)})


CASES = {
    "ransac_easy": case_ransac_easy,
    "ransac_hard": case_ransac_hard,
    "ransac_cap": case_ransac_cap,
    "ransac_degenerate": case_ransac_degenerate,
    "ransac_params": case_ransac_params,
    "ransac_full": case_ransac_full,
    "ransac_13obj_hard": case_ransac_13obj_hard,
    "ransac_1080p": case_ransac_1080p,
    "ls_plain": case_ls_plain,
    "ls_filter": case_ls_filter,
    "ls_full": case_ls_full,
    "ls_grad": case_ls_grad,
    "pose_eval": case_pose_eval,
    "pose_eval_pvnet": case_pose_eval_pvnet,
    "poses_pnp": case_poses_pnp,
    "unmap": case_unmap,
    "pose_metric": case_pose_metric,
}


def main(argv):
    if not os.path.isdir(REFERENCE):
        raise SystemExit("/root/reference is not here: the golden vectors can only be regenerated in the build container")
    tf, mods = load_reference()
    np.seterr(all="ignore")
    for name in argv or list(CASES):
        CASES[name](tf, mods)


if __name__ == "__main__":
    main(sys.argv[1:])
