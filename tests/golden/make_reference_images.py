"""Golden images for the end-to-end render test: the renders the reference SHIPS next to
its scenes (scenes/torus/lmc_timeuse_44.689152s.exr, scenes/veachdoor/lmc_timeuse_30.236183s.exr), box-
filtered by 8 x 8 to small float32 RGB arrays.  Run in the build container (needs /root/reference and
OpenCV's EXR reader); the output tests/golden/reference_images.npz is committed."""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

REF = "/root/reference/scenes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_images.npz")


def load(path, f=8):
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(np.float32)    # BGR -> RGB
    h, w, _ = im.shape
    return im[: h // f * f, : w // f * f].reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3)).astype(np.float32)


np.savez_compressed(OUT,
                    torus_lmc=load(os.path.join(REF, "torus", "lmc_timeuse_44.689152s.exr")),
                    torus_h2mc=load(os.path.join(REF, "torus", "h2mc_timeuse_45.381592s.exr")),
                    door_lmc=load(os.path.join(REF, "veachdoor", "lmc_timeuse_30.236183s.exr")))
print({k: v.shape for k, v in np.load(OUT).items()}, os.path.getsize(OUT))
