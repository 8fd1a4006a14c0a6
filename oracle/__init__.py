"""CPU oracle for the DreamMat SDS hot path.  TEST INFRASTRUCTURE ONLY.

A restatement (torch-CPU / numpy / plain C) of the reference algorithm, function by
function, each citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``dreammat_b200/`` does.

PARITY STATUS: pinned where the reference's own code can run, unpinned where it cannot.
The reference ships no tests, golden vectors or known-answer fixtures
and cannot be imported as a package in the build container (pytorch_lightning / omegaconf /
diffusers / nvdiffrast / tinycudann / envlight / _raytracing are absent).  Its plain-torch
function bodies CAN be executed: tests/golden/make_golden.py lifts them out of
/root/reference by AST and runs them on seeded inputs; tests/test_oracle_golden.py pins this
oracle against the resulting vectors (tests/golden/reference_vectors.pt) for
  a1  collate / cameras / rays / mvp (data/uncond.py:723-821, utils/ops.py:179-292),
  a2  ControlNet normal / depth maps (raytracing_renderer.py:326-343),
  a3  tangent frame + position jitter (raytracing_renderer.py:161-173, 306-316),
  a4  DreamMatMaterial.forward -> shade_raytracing, forward AND autograd backward, with the
      reference's sampling tables, env lookup and occlusion semantics (dreammat_material.py),
  a8/a9  the CSD combination, loss_sds and its gradient (dreammat_guidance.py:440-497, 584-602),
  schedules C(), vertex normals, material export;
and, executed END TO END with the oracle's own pieces standing in for the absent native calls
(tests/golden/make_renderer_golden.py, make_guidance_golden.py, make_splitsum_golden.py),
  a2/a3/a6  the whole RaytraceRender.forward composition (raytracing_renderer.py:110-222) -> render_forward,
  a5        DreamMatMaterial.forward(use_raytracing=False) -> shade_splitsum on the real FG LUT, forward + backward,
  a7-a9     the whole StableDiffusionLightGuidance.__call__ and update_step over dreammat.yaml's schedules
            (dreammat_guidance.py:205-316,388-640) -> sd.guidance_step: loss, logged norms, d loss / d rgb.
PARITY UNPINNED for the arithmetic that lives inside absent native packages -- tiny-cuda-nn
hash grid, nvdiffrast rasterise / antialias / texture, envlight cube maps, the _raytracing BVH,
diffusers UNet / ControlNet / VAE: restated from their published form and pinned only by
``load/lights/bsdf_256_256.bin``, the published parameter counts and self-consistency checks
(brute force vs BVH, autograd vs finite differences, white furnace, MC -> split-sum convergence).
"""
