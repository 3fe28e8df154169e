"""TEST INFRASTRUCTURE ONLY -- runs oracle/_ref/dynaframe_ref, i.e. the reference's OWN path
sources (CDecodeGray.cpp, CDecodePhase.cpp, CCalculation.cpp, CSensorV.cpp, GlobalFunction.cpp)
compiled in place from /root/reference by oracle/Makefile against the minimal OpenCV stand-in in
oracle/ref_shim/, on inputs laid out the way the reference reads them (CSensorV.cpp:35-41,
CCalculation.cpp:86-93,124-127,538).  Only tests/, __graft_entry__.smoke() and bench.py's
reference legs may import this.  The binary is built in the container that has /root/reference
and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "_ref", "dynaframe_ref")
GROUP = os.path.join("20161103", "MoveBoard1103")      # CSensorV.cpp:35


def available() -> bool:
    return os.path.exists(BINARY) and os.access(BINARY, os.X_OK)


def build() -> bool:
    """Compile the reference sources where they lie (needs /root/reference); False if absent."""
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return available()


def _yaml_matrix(name, m) -> str:
    m = np.asarray(m, dtype=np.float64)
    rows, cols = (m.shape if m.ndim == 2 else (m.size, 1))
    vals = ", ".join(f"{v:.17e}" for v in m.reshape(-1))
    return f"{name}: !!opencv-matrix\n   rows: {rows}\n   cols: {cols}\n   dt: d\n   data: [ {vals} ]\n"


class Workspace:
    """A DATA_PATH directory + working directory in the reference's own layout."""

    def __init__(self, cfg, cal, planes: np.ndarray, dyna_frames: np.ndarray | None = None, root: str | None = None):
        from structured_light_calculation_b200 import synth   # BMP writer + layout (host helpers, no kernels)
        self.cfg = cfg
        self.root = root or tempfile.mkdtemp(prefix="dynaframe_ref_")
        self._own = root is None
        self.data = os.path.join(self.root, "data")
        self.cwd = os.path.join(self.root, "cwd")
        self.out = os.path.join(self.root, "out")
        for d in (self.data, os.path.join(self.cwd, "Patterns"), self.out):
            os.makedirs(d, exist_ok=True)
        with open(os.path.join(self.data, "parameters.yml"), "w") as f:          # CCalculation.cpp:86-88,124-131
            f.write("%YAML:1.0\n" + _yaml_matrix("CamMat", cal.cam) + _yaml_matrix("ProMat", cal.pro) +
                    _yaml_matrix("R", cal.R) + _yaml_matrix("T", cal.T))
        with open(os.path.join(self.cwd, "Patterns", "vGrayCode.txt"), "w") as f:   # the format of the reference's table
            for b in range(1 << cfg.gray_digits):
                f.write(f"{b} {b ^ (b >> 1)}\n")
        synth.write_reference_layout(os.path.join(self.data, GROUP), cfg, planes, dyna_frames=dyna_frames)
        self.n_dyna = 0 if dyna_frames is None else int(dyna_frames.shape[0])

    def env(self, maxnum: int) -> dict:
        c = self.cfg
        e = dict(os.environ)
        e.update({
            "DYNAFRAME_PROJECTOR_RESLINE": str(c.projector_width), "DYNAFRAME_CAMERA_RESLINE": str(c.width),
            "DYNAFRAME_CAMERA_RESROW": str(c.height), "DYNAFRAME_GRAY_V_NUMDIGIT": str(c.gray_digits),
            "DYNAFRAME_PHASE_NUMDIGIT": str(c.phase_steps), "DYNAFRAME_DATA_PATH": self.data + "/",
            "DYNAFRAME_MAXNUM": str(maxnum), "DYNAFRAME_FOV_MIN_DISTANCE": str(int(c.fov_min)),
            "DYNAFRAME_FOV_MAX_DISTANCE": str(int(c.fov_max)), "DYNAFRAME_CWD": self.cwd,
        })
        return e

    def _load(self, name, dtype):
        c = self.cfg
        return np.fromfile(os.path.join(self.out, name), dtype).reshape(c.height, c.width)

    def run_first(self) -> dict:
        """Init + FillFirstProjectorU + FillCoordinate(0) of the reference's CCalculation."""
        if self.cfg.phase_steps != 4:
            raise ValueError("the reference reads exactly four phase images (CDecodePhase.cpp:59-62)")
        res = subprocess.run([BINARY, "first", self.out], env=self.env(1), stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0 or "dynaframe_ref ok" not in res.stdout:
            raise RuntimeError(f"dynaframe_ref first failed ({res.returncode}):\n{res.stdout[-2000:]}")
        out = {k: self._load(k + ".f64", np.float64) for k in ("gray", "phase", "projU", "x", "y", "z", "cC", "cD")}
        sc = np.fromfile(os.path.join(self.out, "A_B_P.f64"), np.float64)
        out["A"], out["B"], out["P"] = sc[0], sc[1], sc[2:].reshape(3, 4)
        return out

    def run_full(self) -> dict:
        """Init + CalculateFirst + CalculateOther; also returns the text clouds the reference wrote."""
        n = self.n_dyna
        if n < 2:
            raise ValueError("need at least two dynamic frames")
        res = subprocess.run([BINARY, "full", self.out], env=self.env(n), stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0 or "dynaframe_ref ok" not in res.stdout:
            raise RuntimeError(f"dynaframe_ref full failed ({res.returncode}):\n{res.stdout[-2000:]}")
        frames = []
        for f in range(n):
            d = {"stripB": self._load(f"stripB{f}.f32", np.float32), "stripW": self._load(f"stripW{f}.f32", np.float32),
                 "projU": self._load(f"projU{f}.f64", np.float64)}
            for k in "xyz":
                d[k] = self._load(f"{k}{f}.f64", np.float64)
            if f > 0:
                d["deltaP"] = self._load(f"deltaP{f}.f32", np.float32)
                d["deltaZ"] = self._load(f"deltaZ{f}.f64", np.float64)
            frames.append(d)
        clouds = self._clouds(n)
        return {"frames": frames, "clouds": clouds}

    def _clouds(self, n: int) -> list[bytes]:
        # CCalculation.cpp:192-197,310-315: DATA_PATH + "Res1103\\MoveBoard1103\\" + "PointCloud\\" + name + ".txt";
        # on this platform the backslashes are ordinary file-name characters
        clouds = []
        for f in range(n):
            name = "Res1103\\MoveBoard1103\\PointCloud\\" + ("iFrame" if f == 0 else f"cFrame{f}") + ".txt"
            with open(os.path.join(self.data, name), "rb") as fh:
                clouds.append(fh.read())
        return clouds

    def run_app(self) -> dict:
        """The reference program as main.cpp:42-45 runs it (no plane dumps): wall seconds + its clouds."""
        import time
        n = self.n_dyna
        t0 = time.perf_counter()
        res = subprocess.run([BINARY, "app", self.out], env=self.env(n), stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        secs = time.perf_counter() - t0
        if res.returncode != 0 or "dynaframe_ref ok" not in res.stdout:
            raise RuntimeError(f"dynaframe_ref app failed ({res.returncode}):\n{res.stdout[-2000:]}")
        return {"seconds": secs, "clouds": self._clouds(n)}

    def time_first(self, reps: int) -> list[float]:
        res = subprocess.run([BINARY, "time", str(reps)], env=self.env(1), stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"dynaframe_ref time failed ({res.returncode}):\n{res.stderr[-2000:]}")
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("{\"seconds\"")][-1]
        return json.loads(line)["seconds"]

    def close(self):
        if self._own:
            shutil.rmtree(self.root, ignore_errors=True)
