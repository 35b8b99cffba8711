"""Generates tests/golden/*.  Run in the build container (needs /root/reference, which does NOT
exist on the GPU box):   python tests/golden/make_golden.py

  model_ply_props.npy   the 9 x 62 f32 vertex properties of the reference's only model fixture,
                        /root/reference/coverage/model.ply (test DATA, not source code)
  golden.json           known answers:
     * "survey_kat": values derived in SURVEY.md §8(c) from the reference formulas in mixed
       f64/f32 numpy (NOT reference output; the reference cannot run here) — compared with a
       tolerance of a few ulp of ndc.z;
     * "oracle": outputs of oracle/splat_oracle.c at the commit that generated this file —
       bit-exact regression pins for the oracle itself.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

raw = open("/root/reference/coverage/model.ply", "rb").read()
end = raw.index(b"end_header\n") + len(b"end_header\n")
props = np.frombuffer(raw[end:], dtype="<f4").reshape(9, 62).copy()
np.save(os.path.join(HERE, "model_ply_props.npy"), props)


def hexes(a):
    return [f"0x{int(x):08x}" for x in np.asarray(a).view(np.uint32)]


gold = {"survey_kat": {
    # tests/common/given.rs + tests/e2e/viewer.rs:42-48 scene, SURVEY.md §8(c)
    "single": {"key": "0x3dced7e4", "pixel_centre": [601.42, 600.98], "quad_half_extent": 1536.0,
               "clip": [0.172917, -0.172053, 0.890042, 0.990033]},
    # coverage/model.ply under examples/simple.rs:163-186 defaults at 1280x720
    "model_ply": {"culled": [0, 4, 7], "visible": [1, 2, 3, 5, 6, 8],
                  "keys": {"1": "0x3cccb860", "5": "0x3cccb860", "2": "0x3c8873e0", "3": "0x3c4ca380", "6": "0x3c4ca380", "8": "0x3c4ca380"}},
}}

# ---- oracle pins
g1 = np.zeros(1, dtype=ob.GAUSSIAN_DTYPE)
g1["pos"][0] = (0, 0, 1); g1["rot"][0] = (0, 0, 0, 1); g1["scale"][0] = (1, 1, 1); g1["color"][0] = (255, 0, 0, 255)
m1 = ob.OracleModel(ob.pack_gaussians(g1), 1)
cam1 = ob.camera_pod((0, 0, 0), 0.1, 0.1, 1024, 1024)
gt = ob.gaussian_transform_pod()
p1 = ob.preprocess(m1, cam1, gt)
s1 = ob.project(m1, cam1, gt, p1["indices"][:1])
img1, _ = ob.render(m1, cam1, gt)
gs = ob.gaussians_from_ply_props(props)
mt = ob.model_transform_pod((0, 0, 0), (0, 0, float(np.sin(np.float32(np.pi) / 2)), float(np.cos(np.float32(np.pi) / 2))), (1, 1, 1))
mp = ob.OracleModel(ob.pack_gaussians(gs), 9, model_transform=mt)
camp = ob.camera_pod((0, 0, 0), 0.0, 0.0, 1280, 720)
pp = ob.preprocess(mp, camp, gt)
V = pp["count"]
sk, si = ob.radix_sort(pp["keys"][:V].view(np.uint32), pp["indices"][:V])
imgp, stp = ob.render(mp, camp, gt)
gold["oracle"] = {
    "single": {"camera_pod": hexes(np.frombuffer(bytes(cam1), dtype=np.uint32)), "key": hexes(p1["keys"][:1])[0],
               "draw_args": p1["draw_args"].tolist(), "sort_args": p1["sort_args"].tolist(),
               "splat": {k: float(s1[k][0]) for k in ("cx", "cy", "ax", "ay", "bx", "by", "r", "g", "b", "a", "ext_x", "ext_y")},
               "image_sum_rgba": [int(img1[..., c].astype(np.int64).sum()) for c in range(4)],
               "pixel_600_600": img1[600, 600].tolist(), "pixel_0_0": img1[0, 0].tolist()},
    "model_ply": {"colors": gs["color"].tolist(), "mask": hexes(pp["mask"]), "count": V,
                  "indices_compacted": pp["indices"][:V].tolist(), "keys_compacted": hexes(pp["keys"][:V]),
                  "indices_sorted": si.tolist(), "keys_sorted": hexes(sk),
                  "image_sum_rgba": [int(imgp[..., c].astype(np.int64).sum()) for c in range(4)],
                  "alive_pixels": int(stp["alive_pixels"])},
    "strides": {f"{sh},{cov}": ob.pod_stride(sh, cov) for sh in range(4) for cov in range(3)},
    "exp_neg_poly": {str(x): hexes(np.array([ob.exp_neg_poly(x)], dtype=np.float32))[0] for x in (0.0, 0.5, 1.0, 2.25, 9.0, 50.0, 100.0)},
}
json.dump(gold, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
print("wrote golden.json; model.ply visible:", pp["indices"][:V].tolist(), "sorted:", si.tolist())
