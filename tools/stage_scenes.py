#!/usr/bin/env python3
"""Stage the reference's bundled scenes into scenes/<name>/ of this repo.

Copies the scene XML + mesh + image files verbatim (they are input DATA, not source; the loader decodes PNG / JPEG /
OpenEXR itself) and, as the ground truth the native decoders are tested against, decodes every image with OpenCV into a
tiny raw container tests/golden/decoded/<file>.rawf -- also the hand-over format for images in any other encoding:

    magic 'RAWF' | int32 width | int32 height | int32 is8bit | RGB[h][w][3] (row 0 = top)
    payload is uint8 when is8bit (the loader divides by 255.0f), float32 otherwise

(the reference links OpenImageIO for this, src/image.cpp:5-45, src/bitmaptexture.h:99-146).  `is8bit` drives the 2.2 gamma rule of
src/bitmaptexture.h:136-144.  Run once in the build container:  python tools/stage_scenes.py
"""
import os, shutil, struct, sys
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/scenes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
DECODED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "decoded")


def write_rawf(src, dst):
    img = cv2.imread(src, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("cannot decode " + src)
    is8 = img.dtype == np.uint8
    if img.ndim == 2:
        img = np.stack([img] * 3, axis=-1)
    img = img[..., :3][..., ::-1]  # BGR(A) -> RGB
    h, w = img.shape[:2]
    with open(dst, "wb") as fo:
        fo.write(b"RAWF")
        fo.write(struct.pack("<iii", w, h, 1 if is8 else 0))
        if is8:
            fo.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
        else:
            fo.write(np.ascontiguousarray(img.astype(np.float32), dtype="<f4").tobytes())
    print("  %s -> %s  %dx%d 8bit=%d" % (os.path.basename(src), os.path.basename(dst), w, h, is8))


for scene in ("torus", "veachdoor"):
    sdir = os.path.join(REF, scene)
    odir = os.path.join(OUT, scene)
    os.makedirs(os.path.join(odir, "data"), exist_ok=True)
    for fn in os.listdir(sdir):
        if fn.endswith(".xml"):
            shutil.copyfile(os.path.join(sdir, fn), os.path.join(odir, fn))
    for fn in sorted(os.listdir(os.path.join(sdir, "data"))):
        src = os.path.join(sdir, "data", fn)
        ext = fn.rsplit(".", 1)[-1].lower()
        if ext in ("exr", "png", "jpg", "jpeg"):
            os.makedirs(DECODED, exist_ok=True)
            write_rawf(src, os.path.join(DECODED, fn + ".rawf"))
            shutil.copyfile(src, os.path.join(odir, "data", fn))     # the loader decodes PNG / EXR / JPEG itself (csrc/host/image_decode.h)
        else:
            shutil.copyfile(src, os.path.join(odir, "data", fn))
    print("staged", scene)
