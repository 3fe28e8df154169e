"""B200-native drop-in for DynaFrame's first-frame structured-light path.

Gray-code + N-step phase-shift decode -> temporal unwrap -> projector/camera
triangulation -> float4 XYZ + validity mask, as one fused sm_100a kernel behind
a C ABI (include/slcalc_b200.h).  This Python package is host plumbing only:
the ctypes binding (`capi`), the OpenCV-YAML calibration reader
(`calibration`), workload definitions (`configs`) and the synthetic scene
renderer used by tests and bench (`synth`).  There is no CPU fallback: every
compute entry point raises if the CUDA library is missing.
"""
from .configs import StackConfig, CONFIGS  # noqa: F401

__all__ = ["StackConfig", "CONFIGS"]
__version__ = "0.1.0"
