"""TEST INFRASTRUCTURE ONLY -- CPU restatement of what CSensor::LoadDatas gets from
`imread(path, CV_LOAD_IMAGE_GRAYSCALE)` for a .bmp file (CSensorV.cpp:111-114).

The decoder is third-party (OpenCV highgui, BmpDecoder + utils.cpp; not under /root/reference):
  * rows are stored bottom-up unless biHeight < 0, each padded to 4 bytes;
  * 8 bpp: index -> palette entry -> gray, gray = (B*1868 + G*9617 + R*4899 + 8192) >> 14
    (CvtPaletteToGray; cB/cG/cR = 0.114/0.587/0.299 in 14-bit fixed point), which is the
    identity for a gray palette;
  * 24 / 32 bpp: the same fixed-point formula per pixel (icvCvt_BGR2Gray_8u_C3C1R /
    icvCvt_BGRA2Gray_8u_C4C1R), alpha ignored.
Pinned against cv2.imread of the container's OpenCV 4.13 through tests/golden/bmp_cases.npz.
Only tests/ may import this.
"""
from __future__ import annotations

import struct

import numpy as np

CB, CG, CR, SCALE = 1868, 9617, 4899, 14


def bgr_to_gray(b, g, r):
    return ((b.astype(np.int64) * CB + g.astype(np.int64) * CG + r.astype(np.int64) * CR + (1 << (SCALE - 1)))
            >> SCALE).astype(np.uint8)


def decode_bmp_gray(data: bytes) -> np.ndarray:
    magic, _size, _r1, _r2, off = struct.unpack_from("<2sIHHI", data, 0)
    if magic != b"BM":
        raise ValueError("not a BMP")
    hdr = struct.unpack_from("<I", data, 14)[0]
    if hdr < 40:
        raise ValueError("unsupported header")
    w, h, _planes, bpp, comp, _sz, _xp, _yp, used, _imp = struct.unpack_from("<iiHHIIiiII", data, 18)
    if comp != 0 or bpp not in (8, 24, 32):
        raise ValueError("unsupported BMP flavour")
    top_down = h < 0
    h = abs(h)
    stride = (w * (bpp // 8) + 3) & ~3
    rows = np.frombuffer(data, np.uint8, count=stride * h, offset=off).reshape(h, stride)
    if not top_down:
        rows = rows[::-1]
    if bpp == 8:
        n = used if used else 256
        pal = np.frombuffer(data, np.uint8, count=4 * n, offset=14 + hdr).reshape(n, 4)
        lut = np.zeros(256, np.uint8)
        lut[:n] = bgr_to_gray(pal[:, 0], pal[:, 1], pal[:, 2])
        return lut[rows[:, :w]]
    px = rows[:, : w * (bpp // 8)].reshape(h, w, bpp // 8)
    return bgr_to_gray(px[..., 0], px[..., 1], px[..., 2])
