"""Writes tests/golden/colormap_jet.npy: OpenCV's COLORMAP_JET as a 256 x 3 uint8 table (BGR), i.e.
cv2.applyColorMap of the ramp 0..255.  The table is data of the real OpenCV (cv2 4.13 wheel); the CUDA path, the
oracle and the stand-in applyColorMap of oracle/cvshim all take it as an input, none of them embeds it.

    python tests/golden/make_colormap_golden.py
"""
import os

import cv2
import numpy as np

if __name__ == "__main__":
    ramp = np.arange(256, dtype=np.uint8).reshape(256, 1)
    lut = cv2.applyColorMap(ramp, cv2.COLORMAP_JET).reshape(256, 3)
    np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "colormap_jet.npy"), lut)
    print(lut[:3], lut[127], lut[-3:])
