"""Generates tests/golden/resize_golden.npz by running OpenCV (the reference's own dependency for frame ingest) in THIS
container: small seeded inputs and cv2's outputs for INTER_LINEAR / INTER_NEAREST, incl. up/down scaling, odd sizes,
the exact-2x case and single-channel input.  Run: python oracle/make_resize_golden.py"""
import os

import cv2
import numpy as np

CASES = [(54, 96, 32, 64, 3), (48, 85, 52, 102, 3), (33, 47, 64, 128, 3), (64, 128, 32, 64, 3), (40, 30, 17, 23, 1),
         (27, 48, 16, 32, 3), (108, 192, 64, 128, 3)]

if __name__ == '__main__':
    rng = np.random.default_rng(7)
    out = {'cv2_version': np.array(cv2.__version__)}
    for i, (sh, sw, dh, dw, cn) in enumerate(CASES):
        img = rng.integers(0, 256, size=(sh, sw, cn) if cn > 1 else (sh, sw), dtype=np.uint8)
        lab = rng.integers(0, 21, size=(sh, sw), dtype=np.uint8)
        out['src_%d' % i] = img
        out['lab_%d' % i] = lab
        out['linear_%d' % i] = cv2.resize(img, (dw, dh))
        out['nearest_%d' % i] = cv2.resize(lab, (dw, dh), interpolation=cv2.INTER_NEAREST)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'resize_golden.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')
