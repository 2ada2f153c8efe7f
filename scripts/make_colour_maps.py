"""Writes rustronomy-watershed_b200/data/colour_maps.npz: the 256-entry viridis / magma / plasma / inferno tables
(matplotlib's, CC0) as uint8 RGB, taken from OpenCV's built-in copies.  Run once; the package reads the file."""
import os
import cv2
import numpy as np

out = {}
for name in ("viridis", "magma", "plasma", "inferno"):
    code = getattr(cv2, "COLORMAP_" + name.upper())
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), code).reshape(256, 3)
    out[name] = np.ascontiguousarray(lut[:, ::-1])          # BGR -> RGB
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rustronomy-watershed_b200", "data", "colour_maps.npz")
np.savez_compressed(path, **out)
print(path, {k: v[[0, 128, 255]].tolist() for k, v in out.items()})
