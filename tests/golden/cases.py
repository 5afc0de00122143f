"""Definition of the golden-fixture cases (inputs are regenerated from these seeds; the fixture
stores a SHA-256 of the inputs so that drift in the generator is detected)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _harness as h  # noqa: E402

sm = h.scene_mod

# name -> dict(scene kwargs, view, mode, bg, degree)
CASES = {
    "simple_cloud": dict(kind="simple", W=192, H=96, view=("identity",), mode="sh", bg=(0.0, 0.0, 0.0), degree=0),
    "rand_sh3": dict(kind="rand", P=1500, W=192, H=96, k=0.03, seed=101, view=("identity",), mode="sh",
                     bg=(0.0, 0.0, 0.0), degree=3),
    "rand_odd_white": dict(kind="rand", P=1500, W=205, H=99, k=0.04, seed=102, view=("random", 31), mode="sh",
                           bg=(1.0, 1.0, 1.0), degree=2),
    "rand_colors": dict(kind="rand", P=1200, W=160, H=80, k=0.04, seed=103, view=("random", 32), mode="colors",
                        bg=(0.2, 0.4, 0.6), degree=0),
    "rand_cov": dict(kind="rand", P=1200, W=160, H=80, k=0.04, seed=104, view=("random", 33), mode="cov",
                     bg=(0.0, 0.0, 0.0), degree=1),
    "pole_seam": dict(kind="rand", P=1500, W=256, H=128, k=0.02, seed=105, pole_frac=0.25, seam_frac=0.15,
                      view=("identity",), mode="sh", bg=(0.0, 0.0, 0.0), degree=3),
    # perspective camera (camera_type 1): 90 and 50 degree fov, Gaussians all around the camera so the
    # near plane and the 1.3 tan(fov/2) clamp are exercised; one render_depth frame (forward only)
    "pin_sh3": dict(kind="rand", P=4000, W=192, H=112, k=0.03, seed=106, view=("pinhole", 41, 90.0), mode="sh",
                    bg=(0.0, 0.0, 0.0), degree=3),
    "pin_cov_odd": dict(kind="rand", P=4000, W=205, H=99, k=0.04, seed=107, view=("pinhole", 42, 50.0), mode="cov",
                        bg=(0.3, 0.2, 0.1), degree=1),
    "pin_depth": dict(kind="rand", P=3000, W=160, H=96, k=0.03, seed=108, view=("pinhole", 43, 70.0), mode="sh",
                      bg=(0.0, 0.0, 0.0), degree=2, render_depth=True),
}
GRAD_SEED = 777


def build(name):
    c = CASES[name]
    if c["kind"] == "simple":
        scene = sm.simple_cloud()
        scene.W, scene.H = c["W"], c["H"]
    else:
        scene = sm.make_scene(c["P"], c["W"], c["H"], c["k"], c["seed"], pole_frac=c.get("pole_frac", 0.0),
                              seam_frac=c.get("seam_frac", 0.0), near_frac=0.01)
    if c["view"][0] == "pinhole":
        view = sm.perspective_view(c["view"][1], scene.W, scene.H, c["view"][2])
    else:
        view = sm.identity_view() if c["view"][0] == "identity" else sm.random_view(c["view"][1])
    dL = sm.make_grad_image(scene.W, scene.H, GRAD_SEED)
    return scene, view, dL, c


def input_hash(scene, view, dL):
    m = hashlib.sha256()
    for a in (scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs, view[0], view[1], dL) + tuple(
            np.asarray(v, dtype=np.float32) for v in view[2:]):
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()
