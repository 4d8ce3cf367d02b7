"""Generate tests/golden/aug_crop.npz by running THE REFERENCE's own pylib/HumanAug.crop (oracle/_ref/ref_humanaug.py, streamed
from /root/reference by oracle/make_ref.py, its `scipy.misc` bound to the restated functions of oracle/aug_oracle.py because
SciPy removed them in 1.3) on deterministic synthetic photographs.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):
    python oracle/make_ref.py && python oracle/gen_golden_aug.py
The fixture pins oracle/aug_oracle.py::crop (tests/test_oracle_aug.py) and, through it, the CUDA kernels of
pose_adv_aug_b200/csrc/warp.cu (tests/test_aug_gpu.py).  Per case it stores the parameters, the SHA-256 of the reference's
256 x 256 x 3 uint8 result and a strided sample of it; two cases are stored in full.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_ref, synth          # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (image height, width, photo seed, centre x, centre y, scale, rotation): the branches of crop() --
CASES = [
    (480, 640, 1, 320.0, 240.0, 1.10, 0.0),          # plain crop, window inside the image, up-scaling resize
    (480, 640, 2, 320.0, 240.0, 1.10, 25.0),         # rotation: padded window, PIL rotate, padding removed
    (720, 1280, 3, 600.0, 400.0, 2.70, 0.0),         # scale_factor >= 2: whole image shrunk first (float32 byte-scaling)
    (720, 1280, 4, 600.0, 400.0, 2.70, -31.5),       # shrink + rotation
    (720, 1280, 5, 100.0, 80.0, 1.90, 12.0),         # window hangs over the top-left corner: zero padding joins the min/max
    (500, 375, 6, 360.0, 480.0, 0.60, 0.0),          # small person near the bottom-right corner, strong up-scaling
    (1080, 1920, 7, 1800.0, 1000.0, 3.90, 44.0),     # large shrink factor (wide resampling filter) + rotation + padding
    (480, 640, 8, 10.0, 470.0, 1.30, -15.0),         # centre almost outside
    (600, 800, 9, 400.5, 300.25, 1.28, 0.0),         # scale * 200 == 256: resize by exactly 1 (identity taps)
    (333, 517, 10, 250.0, 160.0, 1.71, 59.9),        # odd sizes, large angle
    (720, 1280, 11, 640.0, 360.0, 2.56, 0.0),        # scale_factor exactly 2.0 (boundary of the shrink branch)
    (720, 1280, 12, 640.0, 360.0, 2.5599, -3.0),     # just below it
]


def case_inputs(k):
    h, w, seed, cx, cy, s, r = CASES[k]
    img = synth.make_photo(h, w, seed)
    return img, np.array([cx, cy], dtype=np.float32), np.array([s], dtype=np.float32), r


def main():
    make_ref.generate(quiet=True)
    R = make_ref.load_humanaug()
    assert R is not None, "reference not found"
    out = {"n_cases": np.int64(len(CASES)), "params": np.array([c for c in CASES], dtype=np.float64)}
    for k in range(len(CASES)):
        img, c, s, r = case_inputs(k)
        res = R.crop(img.copy(), c.copy(), s.copy(), r, 256, 200)
        assert res.shape == (256, 256, 3) and res.dtype == np.uint8
        out["sha%d" % k] = np.frombuffer(hashlib.sha256(res.tobytes()).digest(), dtype=np.uint8)
        out["sample%d" % k] = res[::16, ::16].copy()
        if k in (1, 6):
            out["full%d" % k] = res
        print("case %d: %s mean %.2f sha %s" % (k, CASES[k], res.mean(), hashlib.sha256(res.tobytes()).hexdigest()[:16]))
    np.savez_compressed(os.path.join(GOLD, "aug_crop.npz"), **out)
    print("wrote", os.path.join(GOLD, "aug_crop.npz"), os.path.getsize(os.path.join(GOLD, "aug_crop.npz")), "bytes")


if __name__ == "__main__":
    main()
