"""ctypes wrapper of oracle/ingest_oracle.c + the size / crop arithmetic of the reference's load_images()
(fast3r/dust3r/utils/image.py:68-159).  TEST INFRASTRUCTURE ONLY: the checker for fast3r_b200.ingest."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ingest_oracle.c")
LIB = os.path.join(HERE, "_build", "libingest_oracle.so")
BICUBIC, LANCZOS = 0, 1


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.f3r_oracle_ksize.restype = C.c_int
    return _lib


def resize_plan(w1, h1, size):
    """(new_w, new_h, filter) of _resize_pil_image (image.py:68-75) for the `size != 224` branch of load_images:
    long side -> `size`; LANCZOS when shrinking, BICUBIC otherwise."""
    s = max(w1, h1)
    filt = LANCZOS if s > size else BICUBIC
    return int(round(w1 * size / s)), int(round(h1 * size / s)), filt


def crop_box(w, h, size, square_ok=False):
    """Center crop of load_images (image.py:126-137): (left, top, right, bottom)."""
    cx, cy = w // 2, h // 2
    if size == 224:
        half = min(cx, cy)
        return cx - half, cy - half, cx + half, cy + half
    halfw, halfh = ((2 * cx) // 16) * 8, ((2 * cy) // 16) * 8
    if not square_ok and w == h:
        halfh = 3 * halfw / 4
    return cx - halfw, int(cy - halfh), cx + halfw, int(cy + halfh)


def resize_rgb8(img: np.ndarray, ow: int, oh: int, filt: int) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    out = np.empty((oh, ow, 3), np.uint8)
    lib().f3r_oracle_resize_rgb8(img.ctypes.data_as(C.c_void_p), h, w, out.ctypes.data_as(C.c_void_p), oh, ow, filt)
    return out


def crop_normalize(img: np.ndarray, box) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    l, t, r, b = box
    out = np.empty((3, b - t, r - l), np.float32)
    lib().f3r_oracle_crop_normalize(img.ctypes.data_as(C.c_void_p), h, w, l, t, r - l, b - t, out.ctypes.data_as(C.c_void_p))
    return out


def ingest(img: np.ndarray, size=512, square_ok=False):
    """uint8 RGB (h, w, 3) (already exif-transposed / rotated / cropped to landscape) -> (fp32 (3, H, W), (H, W))."""
    h1, w1, _ = img.shape
    if size == 224:
        nw, nh, filt = resize_plan(w1, h1, round(size * max(w1 / h1, h1 / w1)))
    else:
        nw, nh, filt = resize_plan(w1, h1, size)
    r = resize_rgb8(img, nw, nh, filt)
    box = crop_box(nw, nh, size, square_ok)
    out = crop_normalize(r, box)
    return out, out.shape[1:]
