"""CPU oracle for the DreamMat SDS hot path.  TEST INFRASTRUCTURE ONLY.

A restatement (torch-CPU / numpy / plain C) of the reference algorithm, function by
function, each citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``dreammat_b200/`` does.

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer
fixtures for this path, and it cannot be imported in the build
container (pytorch_lightning / diffusers / nvdiffrast / tinycudann / envlight /
_raytracing are absent).  The only reference-held data the oracle is pinned against is
``load/lights/bsdf_256_256.bin`` (tests/test_oracle_pins.py); everything else is
pinned by self-consistency checks (brute force vs BVH, autograd vs finite
differences, white-furnace, MC -> split-sum convergence).
"""
