"""Calibration input: the OpenCV YAML that CCalculation::Init reads.

Reference: CCalculation.cpp:124-132 (`CamMat`, `ProMat`, `R`, `T`, all f64
`!!opencv-matrix`), fixture `Result.yml`.  OpenCV's FileStorage is not a
dependency here; the small subset of YAML 1.0 it emits is parsed directly.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np


@dataclass
class Calibration:
    cam: np.ndarray  # 3x3 f64 camera intrinsics (CamMat)
    pro: np.ndarray  # 3x3 f64 projector intrinsics (ProMat)
    R: np.ndarray    # 3x3 f64
    T: np.ndarray    # 3 f64

    def scaled(self, cam_scale: float, pro_scale: float) -> "Calibration":
        """Scale intrinsics to another sensor resolution (SURVEY 8d synthetic inputs):
        fu, fv, cu, cv multiply; the pixel-centre convention keeps (c + 0.5)."""
        def sc(K, s):
            K = K.copy()
            K[0, 0] *= s
            K[1, 1] *= s
            K[0, 2] = (K[0, 2] + 0.5) * s - 0.5
            K[1, 2] = (K[1, 2] + 0.5) * s - 0.5
            return K
        return Calibration(sc(self.cam, cam_scale), sc(self.pro, pro_scale), self.R.copy(), self.T.copy())

    def projector_matrix(self) -> np.ndarray:
        """P = ProMat * [R | T] (CCalculation.cpp:141-145), f64 3x4."""
        RT = np.concatenate([self.R, self.T.reshape(3, 1)], axis=1)
        return self.pro @ RT


_MAT_RE = re.compile(
    r"^(\w+):\s*!!opencv-matrix\s*\n\s*rows:\s*(\d+)\s*\n\s*cols:\s*(\d+)\s*\n\s*dt:\s*(\w+)\s*\n\s*data:\s*\[(.*?)\]",
    re.S | re.M,
)


def parse_opencv_yaml(text: str) -> dict:
    out = {}
    for m in _MAT_RE.finditer(text):
        name, rows, cols, dt, data = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)
        if dt not in ("d", "f"):
            raise ValueError(f"{name}: unsupported dt '{dt}'")
        vals = [float(tok) for tok in data.replace("\n", " ").split(",") if tok.strip()]
        if len(vals) != rows * cols:
            raise ValueError(f"{name}: expected {rows * cols} values, got {len(vals)}")
        out[name] = np.array(vals, dtype=np.float64).reshape(rows, cols)
    return out


def load_calibration(path: str) -> Calibration:
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        mats = parse_opencv_yaml(f.read())
    for key in ("CamMat", "ProMat", "R", "T"):
        if key not in mats:
            raise KeyError(f"calibration file {path} lacks '{key}'")
    return Calibration(mats["CamMat"].reshape(3, 3), mats["ProMat"].reshape(3, 3), mats["R"].reshape(3, 3),
                       mats["T"].reshape(3))
