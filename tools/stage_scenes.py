#!/usr/bin/env python3
"""Stage the reference's bundled scenes into scenes/<name>/ of this repo.

Copies the scene XML + mesh files verbatim (they are input DATA, not source) and decodes
every image (EXR envmap, PNG/JPG textures) into a tiny raw float container "<file>.rawf":

    magic 'RAWF' | int32 width | int32 height | int32 is8bit | RGB[h][w][3] (row 0 = top)
    payload is uint8 when is8bit (the loader divides by 255.0f), float32 otherwise

so the C++ host loader needs no image library (the reference links OpenImageIO for this,
src/image.cpp:5-45, src/bitmaptexture.h:99-146).  `is8bit` drives the 2.2 gamma rule of
src/bitmaptexture.h:136-144.  Run once in the build container:  python tools/stage_scenes.py
"""
import os, shutil, struct, sys
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/scenes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")


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
            write_rawf(src, os.path.join(odir, "data", fn + ".rawf"))
            if ext in ("exr", "png"):     # lossless formats are also decoded by the loader itself (csrc/host/image_decode.h)
                shutil.copyfile(src, os.path.join(odir, "data", fn))
        else:
            shutil.copyfile(src, os.path.join(odir, "data", fn))
    print("staged", scene)
