"""Seeded synthetic scenes and views for parity tests and benchmarks.

All randomness comes from numpy's PCG64 (bit-stable across platforms and numpy versions), so the
GPU box, this container and the committed golden fixtures see identical inputs.

Conventions (reference src/gaussian_keyframe.cpp:132-173, cuda_rasterizer/auxiliary.h:85-93,236-248):
camera looks along +z, +x right, +y DOWN (latitude grows with y); ``viewmatrix`` is Tcw transposed,
i.e. flat element [4*c + r] = Tcw[r, c]; ``campos`` is the camera centre in world coordinates;
``rotations`` are unit quaternions (w, x, y, z); ``opacities`` are post-sigmoid, ``scales`` post-exp.
"""
import dataclasses
import math

import numpy as np

BASE_SEED = 20240403

# name -> (P, W, H, k, extras)   k = scale factor relative to distance (calibrated in SURVEY §8(d))
CONFIGS = {
    "C1": dict(P=100_000, W=1024, H=512, k=0.02),
    "C2": dict(P=1_000_000, W=2048, H=1024, k=0.01),
    "C3": dict(P=3_000_000, W=1920, H=960, k=0.005),
    "C4": dict(P=5_000_000, W=7680, H=3840, k=0.0015),
    "C5": dict(P=10_000_000, W=4096, H=2048, k=0.001, pole_frac=0.10, seam_frac=0.02),
}
CONFIG_INDEX = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}


@dataclasses.dataclass
class Scene:
    means3D: np.ndarray     # [P,3] float32
    scales: np.ndarray      # [P,3] float32 (post-exp)
    rotations: np.ndarray   # [P,4] float32 unit (w,x,y,z)
    opacities: np.ndarray   # [P,1] float32 in [0.05,1)
    shs: np.ndarray         # [P,16,3] float32
    W: int
    H: int

    @property
    def P(self):
        return int(self.means3D.shape[0])


def make_scene(P, W, H, k, seed, pole_frac=0.0, seam_frac=0.0, near_frac=0.001, sh_coeffs=16):
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.standard_normal((P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-12
    # polar cap: lat = +-(90deg - |N(0, 8deg)|), uniform longitude
    n_pole = int(round(P * pole_frac))
    if n_pole:
        lat = (math.pi / 2 - np.abs(rng.normal(0.0, math.radians(8.0), n_pole))) * rng.choice([-1.0, 1.0], n_pole)
        lon = rng.uniform(-math.pi, math.pi, n_pole)
        d[:n_pole] = np.stack([np.cos(lat) * np.sin(lon), np.sin(lat), np.cos(lat) * np.cos(lon)], axis=1)
    # seam: within +-1 tile (16 px) of lon = +-pi
    n_seam = int(round(P * seam_frac))
    if n_seam:
        dlon = rng.uniform(-16.0, 16.0, n_seam) * (2 * math.pi / W)
        lon = np.where(dlon >= 0, -math.pi + dlon, math.pi + dlon)
        lat = np.arcsin(rng.uniform(-1.0, 1.0, n_seam))
        d[n_pole:n_pole + n_seam] = np.stack([np.cos(lat) * np.sin(lon), np.sin(lat), np.cos(lat) * np.cos(lon)], axis=1)
    r = np.exp(rng.uniform(math.log(0.5), math.log(20.0), P))
    n_near = int(round(P * near_frac))
    if n_near:
        r[P - n_near:] = rng.uniform(0.01, 0.25, n_near)  # exercises the r^2 <= 0.04 cull
    means = d * r[:, None]
    scales = np.exp(rng.normal(np.log(k * r)[:, None], 0.7, (P, 3)))
    q = rng.standard_normal((P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True) + 1e-12
    opac = rng.uniform(0.05, 1.0, (P, 1))
    shs = np.empty((P, sh_coeffs, 3))
    shs[:, 0, :] = rng.uniform(-1.0, 1.5, (P, 3))
    if sh_coeffs > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.1, (P, sh_coeffs - 1, 3))
    f = np.float32
    return Scene(means.astype(f), scales.astype(f), q.astype(f), opac.astype(f), shs.astype(f), W, H)


def make_config_scene(name, P=None):
    """The SURVEY §8(d) scene for config C1..C5 (optionally with a reduced Gaussian count)."""
    cfg = dict(CONFIGS[name])
    if P is not None:
        cfg["P"] = int(P)
    return make_scene(seed=BASE_SEED + CONFIG_INDEX[name], **cfg)


def identity_view():
    """Camera at the origin, identity pose: (viewmatrix[4,4] = Tcw^T, campos[3])."""
    return np.eye(4, dtype=np.float32), np.zeros(3, dtype=np.float32)


def random_view(seed, ball=0.3):
    """A random SE(3) pose whose centre lies within `ball` of the origin (egocentric capture)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    c = rng.standard_normal(3)
    c *= ball * rng.uniform() ** (1 / 3) / np.linalg.norm(c)
    Tcw = np.eye(4)
    Tcw[:3, :3] = R
    Tcw[:3, 3] = -R @ c
    return Tcw.T.astype(np.float32).copy(), c.astype(np.float32)


def perspective_view(seed, W, H, fovx_deg=90.0, znear=0.01, zfar=100.0, ball=0.3):
    """A random pinhole camera (camera_type 1): (viewmatrix = Tcw^T, projmatrix = full transform, campos,
    tan_fovx, tan_fovy).  The projection is the reference's getProjectionMatrix (3DGS convention: z in [0,1],
    w = z_cam) and projmatrix = viewmatrix @ P^T in row-major terms, i.e. flat element [4*c + r] = (P Tcw)[r, c]
    (reference src/gaussian_keyframe.cpp:142-160)."""
    V, c = random_view(seed, ball)
    tan_fovx = math.tan(math.radians(fovx_deg) / 2)
    tan_fovy = tan_fovx * H / W
    top, right = tan_fovy * znear, tan_fovx * znear
    P = np.zeros((4, 4))
    P[0, 0] = znear / right
    P[1, 1] = znear / top
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    full = (V.astype(np.float64) @ P.T).astype(np.float32)
    return V, np.ascontiguousarray(full), c, np.float32(tan_fovx).item(), np.float32(tan_fovy).item()


def yaw_view(angle):
    """Camera at the origin rotated by `angle` about the vertical (y) axis."""
    c, s = math.cos(angle), math.sin(angle)
    Tcw = np.eye(4)
    Tcw[:3, :3] = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return Tcw.T.astype(np.float32).copy(), np.zeros(3, dtype=np.float32)


def make_grad_image(W, H, seed):
    """Seeded upstream gradient dL/d(out_color) [3,H,W] = N(0,1)/N."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.standard_normal((3, H, W)) / (W * H)).astype(np.float32)


def simple_cloud():
    """The reference's only known-input scene (examples/simple_cloud.cpp:133-166,223-227): three
    Gaussians at (d,-5d,d), (-d,0.5d,-0.7d), (d,d,-d) with d = 1, coloured R/G/B through the SH DC
    term, log-scale -0.3, opacity logit 5, identity rotation and pose.  Returned at a reduced
    default resolution by the callers that render it."""
    d = 1.0
    means = np.array([[d, -5 * d, d], [-d, 0.5 * d, -0.7 * d], [d, d, -d]], dtype=np.float32)
    scales = np.full((3, 3), math.exp(-0.3), dtype=np.float32)
    rot = np.tile(np.array([1, 0, 0, 0], dtype=np.float32), (3, 1))
    opac = np.full((3, 1), 1.0 / (1.0 + math.exp(-5.0)), dtype=np.float32)
    shs = np.zeros((3, 16, 3), dtype=np.float32)
    c0 = 0.28209479177387814
    for i in range(3):
        rgb = np.zeros(3)
        rgb[i] = 1.0
        shs[i, 0, :] = (rgb - 0.5) / c0
    return Scene(means, scales, rot, opac, shs, 2000, 1000)
