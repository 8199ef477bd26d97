"""Synthetic stand-ins for the artefacts the reference ships outside the repo (weights: `.MISSING_LARGE_BLOBS`) or
needs from a camera (frames, teacher label maps).  Used by bench.py and the examples; deterministic given a seed.
The checkpoint has the reference layout {'<tf variable name>:0': float32 ndarray} (utils/utils.py:26-27)."""
from collections import OrderedDict

import numpy as np

from .student import load_graph_spec


def synthetic_checkpoint(tag='cityscapes', seed=1):
    """He-normal kernels; BN gamma~U(0.5,1) / beta~N(1,0.2) on ReLU(6) layers (about 10 % of units clipped, so the
    stack does not amplify rounding noise exponentially the way a zero-centred random BN/ReLU stack does);
    gamma~U(0.5,1.5) / beta~N(0,0.1) on the linear bottlenecks; moving statistics N(0,0.1) / U(0.5,1.5)."""
    spec = load_graph_spec(tag)
    rng = np.random.default_rng(seed)
    linear = set()
    for c in spec['convs']:
        if c['bn'] is not None and c['act'] is None:
            linear.update((c['bn']['beta'], c['bn']['gamma']))
    out = OrderedDict()
    for v in spec['variables']:
        name, shape = v['name'], tuple(v['shape'])
        if name.endswith('weights:0'):
            fan_in = shape[0] * shape[1] * (1 if 'depthwise' in name else shape[2])
            a = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        elif name.endswith('gamma:0'):
            a = rng.uniform(0.5, 1.5, size=shape) if name in linear else rng.uniform(0.5, 1.0, size=shape)
        elif name.endswith('beta:0'):
            a = rng.normal(0.0, 0.1, size=shape) if name in linear else rng.normal(1.0, 0.2, size=shape)
        elif name.endswith('moving_variance:0'):
            a = rng.uniform(0.5, 1.5, size=shape)
        else:                                   # moving_mean, logits biases
            a = rng.normal(0.0, 0.1, size=shape)
        out[name] = a.astype(np.float32)
    return out


def synthetic_frames(n, h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def synthetic_labels(n, h, w, seed=0, block=32, ignore_frac=0.05, num_ids=19):
    """Piecewise-constant teacher label maps (block x block cells), ignore_frac of the pixels set to 255."""
    rng = np.random.default_rng(seed + 1000)
    coarse = rng.integers(0, num_ids, size=(n, -(-h // block), -(-w // block)), dtype=np.uint8)
    lab = np.repeat(np.repeat(coarse, block, axis=1), block, axis=2)[:, :h, :w].copy()
    lab[rng.random(size=(n, h, w)) < ignore_frac] = 255
    return lab
