"""Deterministic synthetic inputs: calibration + scene -> ground truth -> u8 pattern stack.

Pattern *generation* is not in the reference; the conventions here are the
ones its decoders imply:

* Gray pairs: plane 2b is the pattern for bit b of gray(floor(U/gp)), plane
  2b+1 its inverse; b = 0 is the LSB (CDecodeGray.cpp:159,193-198).
* Phase images: I_k = a + b*sin(2*pi*(U-0.5)/T + 2*pi*k/N)
  (CDecodePhase.cpp:59-62 comments give the k*pi/2 steps for N = 4; the -0.5 is
  the pixel-centre offset that CDecodePhase.cpp:70 adds back).
* Scene (camera frame, reference units; working range z in [10,100]): a tilted
  plane at z~60, a sphere (r = 8, centre z~40), a 3-unit step across a vertical
  edge, a far background band beyond FOV_MAX, and a low-albedo patch.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .calibration import Calibration
from .configs import StackConfig


@dataclass
class Scene:
    z: np.ndarray        # [H,W] f64 true depth along the camera ray (z component)
    xyz: np.ndarray      # [H,W,3] f64
    U: np.ndarray        # [H,W] f64 true projector column
    lit: np.ndarray      # [H,W] bool: projector column inside [0, PW)
    albedo: np.ndarray   # [H,W] f64 in (0,1]
    V: np.ndarray = None  # [H,W] f64 true projector row (row 1 of P; the reference never decodes it)


def make_scene(cfg: StackConfig, cal: Calibration, plane_z: float = 60.0) -> Scene:
    H, W = cfg.height, cfg.width
    fu, fv, cu, cv = cal.cam[0, 0], cal.cam[1, 1], cal.cam[0, 2], cal.cam[1, 2]
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    dx = (u - cu) / fu
    dy = (v - cv) / fv
    # tilted plane n.X = d0 with n = (0.06, -0.04, 1): z = d0 / (n . d)
    z = plane_z / (0.06 * dx - 0.04 * dy + 1.0)
    z = np.broadcast_to(z, (H, W)).copy()
    # step: everything right of 62% width is 3 units closer
    z[:, int(0.62 * W):] -= 3.0
    # sphere centred on the ray through (0.35W, 0.45H) at z = 40, radius 8
    cdir = np.array([(0.35 * W - cu) / fu, (0.45 * H - cv) / fv, 1.0])
    centre = 40.0 * cdir
    r = 8.0
    d = np.stack([np.broadcast_to(dx, (H, W)), np.broadcast_to(dy, (H, W)), np.ones((H, W))], axis=-1)
    dd = (d * d).sum(-1)
    dc = d @ centre
    disc = dc * dc - dd * (centre @ centre - r * r)
    hit = disc > 0
    t = np.where(hit, (dc - np.sqrt(np.where(hit, disc, 0.0))) / dd, np.inf)
    z = np.where(hit & (t < z), t, z)
    # far background band along the top edge (beyond FOV_MAX)
    z[: max(1, H // 16), :] = 150.0
    xyz = d * z[..., None]
    P = cal.projector_matrix()
    Xh = np.concatenate([xyz, np.ones((H, W, 1))], axis=-1)
    num = Xh @ P[0]
    den = Xh @ P[2]
    U = num / den
    V = (Xh @ P[1]) / den
    lit = (U >= 0.0) & (U < cfg.projector_width)
    # smooth albedo field in [0.25, 1] plus one low-albedo patch
    yy = v / H
    xx = u / W
    albedo = 0.625 + 0.375 * np.sin(2 * np.pi * (3 * xx + 0.3)) * np.cos(2 * np.pi * (2 * yy + 0.1))
    albedo = np.broadcast_to(albedo, (H, W)).copy()
    albedo[int(0.70 * H): int(0.80 * H), int(0.10 * W): int(0.25 * W)] = 0.02
    return Scene(z=z, xyz=xyz, U=U, lit=lit, albedo=albedo, V=V)


def gray_code(n: np.ndarray) -> np.ndarray:
    return n ^ (n >> 1)


def render_stack(cfg: StackConfig, scene: Scene, noise_sigma: float = 1.0, seed: int = 1234,
                 ambient: float = 6.0, hi: float = 235.0, amp: float = 105.0, horizontal: bool = False) -> np.ndarray:
    """Return the plane-major stack [2G+N][H][W] u8 for one frame set.  horizontal=True renders the
    patterns along the projector ROWS instead (cfg.projector_width is then the projector height and
    cfg.gray_digits the horizontal digit count, what SetNumDigit(n, false) selects)."""
    H, W, G, N = cfg.height, cfg.width, cfg.gray_digits, cfg.phase_steps
    gp, T = cfg.gray_period, cfg.phase_period
    rng = np.random.Generator(np.random.PCG64(seed))
    planes = np.empty((cfg.planes, H, W), dtype=np.uint8)
    U = scene.V if horizontal else scene.U
    alb = scene.albedo
    lit = (scene.lit & (U >= 0.0) & (U < cfg.projector_width)) if horizontal else scene.lit
    kbin = np.clip(np.floor(U / gp), 0, (1 << G) - 1).astype(np.int64)
    g = gray_code(kbin)

    def finish(img):
        if noise_sigma > 0:
            img = img + rng.standard_normal(img.shape) * noise_sigma
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    for b in range(G):
        bit = ((g >> b) & 1).astype(bool) & lit
        on = ambient + alb * hi
        off = np.full_like(on, ambient)
        planes[2 * b] = finish(np.where(bit, on, off))
        planes[2 * b + 1] = finish(np.where(bit | ~lit, off, on))
    theta = 2.0 * np.pi * (U - 0.5) / T
    for k in range(N):
        img = ambient + alb * (hi / 2 + amp * np.sin(theta + 2.0 * np.pi * k / N))
        img = np.where(lit, img, ambient)
        planes[2 * G + k] = finish(img)
    return planes


def synthetic_calibration(cfg: StackConfig, base: Calibration) -> Calibration:
    """Result.yml (a 640x512 camera, 1280-wide projector) scaled to cfg."""
    return base.scaled(cfg.width / 640.0, cfg.projector_width / 1280.0)


def render_dyna_frames(cfg: StackConfig, cal: Calibration, n_frames: int, stripe_period: float = 20.0,
                       z_step: float = 0.05, noise_sigma: float = 1.0, seed: int = 4321) -> np.ndarray:
    """The dynamic-mode input (`dynaCam{i}.bmp` in the reference): one camera image per frame
    of a static sinusoidal stripe pattern while the scene's plane recedes by z_step per frame."""
    rng = np.random.Generator(np.random.PCG64(seed))
    frames = np.empty((n_frames, cfg.height, cfg.width), np.uint8)
    for f in range(n_frames):
        sc = make_scene(cfg, cal, plane_z=60.0 + z_step * f)
        img = 6.0 + sc.albedo * (117.0 + 105.0 * np.sin(2.0 * np.pi * sc.U / stripe_period))
        img = np.where(sc.lit, img, 6.0) + rng.standard_normal(img.shape) * noise_sigma
        frames[f] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return frames


def encode_bmp(pixels: np.ndarray, bpp: int = 8, top_down: bool = False) -> bytes:
    """A BITMAPINFOHEADER .bmp of a [H][W] u8 image: 8 bpp with the gray palette a camera SDK
    writes, or 24 bpp (B = G = R).  Test / bench input for the ingest path (the reference reads
    vGrayCam{i}.bmp etc., CSensorV.cpp:111-114)."""
    import struct
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    H, W = pixels.shape
    if bpp == 24:
        pixels = np.repeat(pixels[:, :, None], 3, axis=2)
    elif bpp != 8:
        raise ValueError("bpp must be 8 or 24")
    row = W * (bpp // 8)
    stride = (row + 3) & ~3
    rows = pixels if top_down else pixels[::-1]
    body = np.zeros((H, stride), np.uint8)
    body[:, :row] = rows.reshape(H, row)
    pal = b""
    if bpp == 8:
        g = np.arange(256, dtype=np.uint8)
        pal = np.stack([g, g, g, np.zeros(256, np.uint8)], axis=1).tobytes()
    off = 14 + 40 + len(pal)
    head = struct.pack("<2sIHHI", b"BM", off + body.size, 0, 0, off)
    info = struct.pack("<IiiHHIIiiII", 40, W, -H if top_down else H, 1, bpp, 0, body.size, 2835, 2835, 0, 0)
    return head + info + pal + body.tobytes()


def write_reference_layout(group_dir: str, cfg: StackConfig, planes: np.ndarray, dyna_frames=None, bpp: int = 8):
    """Write a stack (and optionally a dynamic sequence) as the reference's file layout
    (CSensorV.cpp:35-41): iFrame/vGrayCam{i}.bmp, iFrame/vPhaseCam{i}.bmp, cFrame/dynaCam{i}.bmp.
    Returns the list of first-frame paths in stack-plane order."""
    import os
    G2 = 2 * cfg.gray_digits
    os.makedirs(os.path.join(group_dir, "iFrame"), exist_ok=True)
    paths = []
    for i in range(planes.shape[0]):
        name = f"vGrayCam{i}.bmp" if i < G2 else f"vPhaseCam{i - G2}.bmp"
        path = os.path.join(group_dir, "iFrame", name)
        with open(path, "wb") as f:
            f.write(encode_bmp(planes[i], bpp))
        paths.append(path)
    if dyna_frames is not None:
        os.makedirs(os.path.join(group_dir, "cFrame"), exist_ok=True)
        for i in range(dyna_frames.shape[0]):
            with open(os.path.join(group_dir, "cFrame", f"dynaCam{i}.bmp"), "wb") as f:
                f.write(encode_bmp(dyna_frames[i], bpp))
    return paths
