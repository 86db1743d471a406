"""The torch-CPU twin (timed as the CPU baseline) must agree with the numpy oracle: bit-identical vote counts."""
import numpy as np
import torch

from casapose_b200 import synthetic
from oracle import ransac_voting_np as O
from oracle import ransac_voting_torch as T


def test_twin_matches_numpy_oracle_bit_for_bit():
    for variant in ("easy", "hard"):
        d = synthetic.make_frames(1, 96, 128, (1, 5, 6), variant=variant)
        ref, dbg = O.ransac_voting_layer_all_masks(d["mask"], d["vertex"], 48, seed=5, max_iter=4, return_debug=True)
        out, infos = T.ransac_voting_layer_all_masks(torch.from_numpy(d["mask"]), torch.from_numpy(d["vertex"]), 48,
                                                     seed=5, max_iter=4, return_info=True)
        for c in range(3):
            r, t = dbg[0][c], infos[c]
            assert r["tn"] == t["tn"] and r["rounds"] == t["rounds"]
            for k in range(r["rounds"]):
                assert np.array_equal(r["counts"][k], t["counts"][k].numpy())
        assert np.abs(out.numpy() - ref).max() < 1e-3
