"""Workload geometry: the reference's compile-time constants as a runtime struct.

Reference: StaticParameters.cpp:4-9,16-18,34-35 (resolutions, digit counts,
FOV limits).  CONFIGS holds the five BASELINE.json configurations as read in
SURVEY.md section 8(d).
"""
from __future__ import annotations

from dataclasses import dataclass, replace


@dataclass(frozen=True)
class StackConfig:
    width: int                 # CAMERA_RESLINE
    height: int                # CAMERA_RESROW
    projector_width: int       # PROJECTOR_RESLINE
    gray_digits: int           # GRAY_V_NUMDIGIT: pattern/inverse PAIRS incl. the half-period LSB
    phase_steps: int           # PHASE_NUMDIGIT
    fov_min: float = 10.0      # FOV_MIN_DISTANCE
    fov_max: float = 100.0     # FOV_MAX_DISTANCE
    modulation_min: float = 0.0  # [EXT] b_min in grey levels; 0 = disabled (reference-faithful)
    name: str = ""

    @property
    def planes(self) -> int:
        return 2 * self.gray_digits + self.phase_steps

    @property
    def pixels(self) -> int:
        return self.width * self.height

    @property
    def gray_period(self) -> int:
        """gp = PW / 2^G (CDecodeGray.cpp:183, integer division)."""
        return self.projector_width // (1 << self.gray_digits)

    @property
    def phase_period(self) -> int:
        """T = PW / 2^(G-1) (CCalculation.cpp:550, integer division)."""
        return self.projector_width // (1 << (self.gray_digits - 1))

    @property
    def algorithmic_bytes_per_pixel(self) -> int:
        """u8 planes read once + float4 XYZ + u8 mask written once."""
        return self.planes + 16 + 1

    @property
    def stack_bytes(self) -> int:
        return self.planes * self.pixels

    def with_(self, **kw) -> "StackConfig":
        return replace(self, **kw)


CONFIGS = {
    # the reference's own compile-time defaults (StaticParameters.cpp)
    "reference_default": StackConfig(1280, 1024, 1280, 6, 4, name="reference_default"),
    # BASELINE.json configs[0..4]
    "config1": StackConfig(1280, 1024, 1280, 7, 4, name="config1"),
    "config2": StackConfig(1920, 1200, 2560, 9, 4, name="config2"),
    "config3": StackConfig(2448, 2048, 2048, 8, 8, modulation_min=8.0, name="config3"),
    "config4": StackConfig(1920, 1200, 2560, 9, 4, name="config4"),  # 4096 frame sets of config2
    "config5": StackConfig(4096, 3000, 4096, 10, 12, name="config5"),
}
