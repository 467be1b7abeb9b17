"""CPU oracle for the render-and-compare refinement + multiview matching path.

TEST INFRASTRUCTURE.  Only tests/, `__graft_entry__.smoke()` and bench.py's
`cpu_baseline` / `--impl reference` legs may import this package, and only as the
checker (or the timed CPU baseline) - never the product path.  The product
(`cosypose_b200`) fails loudly when its CUDA library is missing; it has no CPU fallback.

Parity pinning: the reference ships no tests, golden vectors
or fixtures for this path.  The oracle is therefore pinned against outputs of the
reference itself, imported unmodified from /root/reference in the build container
by tests/golden/make_golden.py; the resulting vectors are committed under
tests/golden/*.npz and `tests/test_oracle_golden.py` checks the oracle against them.
The integer RANSAC helpers are checked against the reference's own C++ extension
compiled into oracle/_ref (oracle/Makefile).

Modules
  pose_oracle       plain torch-CPU fp32 restatement of the single-view path
                    (geometry, RoI crop, EfficientNet-B3 trunk, pose head/update)
  multiview_oracle  restatement of candidate matching (symmetric distances, RANSAC
                    models/scoring, scene-level matching) and bundle adjustment
  cext_oracle       pure-Python restatement of the four cosypose_cext functions
"""
