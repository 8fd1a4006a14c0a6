"""Fixture generator for the split-sum (a5) parity tests: REAL assets of the reference, decoded by the reference's own code.

    fg_lut          load/lights/bsdf_256_256.bin read exactly as dreammat_material.py:405-410 does (float32 [1,256,256,2])
    envmap_64x128   load/lights/envmap/map1/map1.exr decoded by the reference's `load_hdr_image` (lifted by AST from
                    dreammat_material.py:65-68), then area-averaged 32x to 64x128 so it fits a fixture (the full map is 100 MB)

Run where /root/reference exists:  python tests/golden/make_splitsum_assets.py   -> tests/golden/splitsum_assets.pt
"""
import ast
import os
import sys

import numpy as np
import torch

REF = "/root/reference/threestudio_dreammat"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitsum_assets.pt")


def main():
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    src = open(os.path.join(REF, "threestudio/models/materials/dreammat_material.py")).read()
    tree = ast.parse(src)
    ns = {"cv2": cv2, "np": np, "torch": torch}
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "load_hdr_image"]
    assert len(fn) == 1
    exec(compile(ast.Module(body=fn, type_ignores=[]), "load_hdr_image", "exec"), ns)
    # the FG_LUT statement of configure() (:405-410), evaluated from the reference source with the path made absolute
    lut_stmt = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "FG_LUT"]
    assert len(lut_stmt) == 1
    code = ast.unparse(lut_stmt[0]).replace('"load/lights/bsdf_256_256.bin"', repr(os.path.join(REF, "load/lights/bsdf_256_256.bin")))
    code = code.replace("'load/lights/bsdf_256_256.bin'", repr(os.path.join(REF, "load/lights/bsdf_256_256.bin")))
    exec(code, ns)
    lut = ns["FG_LUT"].clone()
    assert tuple(lut.shape) == (1, 256, 256, 2) and lut.dtype == torch.float32
    img = torch.tensor(ns["load_hdr_image"](os.path.join(REF, "load/lights/envmap/map1/map1.exr")), dtype=torch.float32)
    small = cv2.resize(img.numpy(), (128, 64), interpolation=cv2.INTER_AREA)
    torch.save({"fg_lut": lut, "envmap_64x128": torch.from_numpy(np.ascontiguousarray(small)),
                "envmap_full_shape": tuple(img.shape), "lut_spot": {"0,0": lut[0, 0, 0].tolist(), "128,128": lut[0, 128, 128].tolist()}}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; map", tuple(img.shape), "->", small.shape, "lut[0,0] =", lut[0, 0, 0].tolist())


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    main()
