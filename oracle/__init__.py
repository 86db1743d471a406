"""CPU oracle for the CASAPose keypoint-voting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU arm.

HOW IT IS PINNED.  The reference (/root/reference, fraunhoferhhi/casapose) is pure
Python on TensorFlow 2.9.1 + tensorflow-addons 0.17.0.  Neither is installed in this
image (nor on the GPU box) and the reference ships no tests, golden vectors or
fixtures for this path.  What pins the restatement instead:
  * tests/golden/*.npz — outputs of the reference's OWN SOURCE FILES
    (casapose/pose_estimation/{ransac_voting,voting_layers_2d,pose_evaluation,
    bpnp_layers}.py, imported unmodified from /root/reference) executed over a numpy
    stand-in for the TensorFlow API (oracle/tf_standin/, generator oracle/make_golden.py):
    vote counts of every round, round counts, keypoints, LS-layer outputs, poses and
    ADD / ADD-S / 2-D verdicts, and the LS layer's gradient (torch.autograd over the
    same source on oracle/tf_standin_torch), for 15 cases incl. BASELINE configs 1, 2, 3 and 5 at full
    size.  tests/test_golden_oracle.py: vote counts bit-identical, keypoints <= 1e-3 px,
    verdicts identical.  This anchors the op order, axis conventions, flips, gates, tie
    rules and control flow to the reference's code;
  * what stays UNPINNED is TensorFlow's own kernels (Eigen's float32 summation order and
    powf, LAPACK-style svd/inv/pinv, tfa's connected components): no TensorFlow can run
    here, the stand-in uses numpy's / scipy's, and DESIGN.md section 3 lists each choice;
  * a published known-answer vector for the Philox4x32-10 generator (Random123),
  * hand-computed known-answer cases for every degenerate branch,
  * a second, independently written torch-CPU twin that must agree bit-for-bit on
    vote counts (``oracle/ransac_voting_torch.py``).
"""
