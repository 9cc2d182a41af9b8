#!/usr/bin/env python3
"""Golden fixture from the reference's OWN render of its default scene.

  /root/reference/docs/semesterbild.png (800x600, rendered by jackra1n/raytracer-rust from data/scenes/semesterbild.json;
  revision and spp unknown)  ->  tests/golden/semesterbild_ref_blocks.npy

The PNG is reduced to 8x8-pixel block means (100x75x3 float32, 8-bit display levels after the renderer's sqrt "gamma"),
which removes the per-pixel Monte-Carlo noise of the unknown sample count and keeps the fixture small.  The GPU box has
no /root/reference, so tests read only the committed .npy.

  python tools/make_golden.py [/root/reference]
"""
import os
import sys

import numpy as np
from PIL import Image

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
img = np.asarray(Image.open(os.path.join(ref, "docs", "semesterbild.png")).convert("RGB")).astype(np.float32)
assert img.shape == (600, 800, 3)
blocks = img.reshape(75, 8, 100, 8, 3).mean(axis=(1, 3)).astype(np.float32)
os.makedirs(root, exist_ok=True)
np.save(os.path.join(root, "semesterbild_ref_blocks.npy"), blocks)
np.save(os.path.join(root, "semesterbild_ref_corner.npy"), img[:4, :4].astype(np.uint8))
print("wrote", blocks.shape, "top-left pixel", img[0, 0])
