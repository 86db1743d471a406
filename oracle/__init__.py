"""CPU oracle for the CASAPose keypoint-voting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU arm.

PARITY UNPINNED.  The reference (/root/reference, fraunhoferhhi/casapose) is pure
Python on TensorFlow 2.9.1 + tensorflow-addons 0.17.0.  Neither is installed in this
image (nor on the GPU box) and the reference ships no tests, golden vectors or
fixtures for this path, so the restatement below cannot be checked against outputs
of the reference itself.  It is an op-by-op restatement (one float32 rounding per
TensorFlow op, no fusion) that cites the reference file:line for every function, and
it is pinned by:
  * a published known-answer vector for the Philox4x32-10 generator (Random123),
  * hand-computed known-answer cases for every degenerate branch,
  * a second, independently written torch-CPU twin that must agree bit-for-bit on
    vote counts (``oracle/ransac_voting_torch.py``).
"""
