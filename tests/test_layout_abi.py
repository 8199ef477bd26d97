"""CPU tests of the boundary: the shared library loads, exports every symbol include/ams_b200.h declares, its
restated topology equals what the reference's model.meta contains (ams_b200/graphs/*.json), and it refuses to run
without a GPU (no CPU fallback).  No compute calls are made."""
import json
import os
import re

import numpy as np
import pytest

from ams_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'ams_b200.h')).read()
    declared = set(re.findall(r'\b(ams_[a-z0-9_]+)\s*\(', header)) - {'ams_net', 'ams_config'}
    lib = nat.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(nat.exported_symbols()), declared ^ set(nat.exported_symbols())
    assert lib.ams_abi_version() == 1


@pytest.mark.parametrize('nc,variant,tag', [(19, 0, 'cityscapes'), (21, 1, 'pascalvoc2012')])
def test_topology_matches_model_meta(nc, variant, tag):
    V, L = nat.layout(nc, variant)
    spec = json.load(open(os.path.join(ROOT, 'ams_b200', 'graphs', tag + '.json')))
    assert [v['name'] for v in V] == [v['name'] for v in spec['variables']]
    assert [v['shape'] for v in V] == [v['shape'] for v in spec['variables']]
    assert [v['name'] for v in V if v['trainable']] == [v['name'] for v in spec['trainable_variables']]
    off = 0
    for v in V:                                              # the trainable arena is dense, in trainable order
        if v['trainable']:
            assert v['offset'] == off
            off += int(np.prod(v['shape']))
    assert off == {19: 2113043, 21: 2113557}[nc]
    assert len(L) == len(spec['convs']) == 55
    for l, c in zip(L, spec['convs']):
        assert l['name'] == c['name']
        assert (l['cin'], l['cout'], l['stride'], l['dilation']) == (c['cin'], c['cout'], c['stride'], c['dilation'])
        assert l['act'] == {None: 0, 'relu': 1, 'relu6': 2}[c['act']]
        if c['bn']:
            assert np.float32(l['eps']) == np.float32(c['bn']['eps'])
            assert np.float32(l['one_minus_decay']) == np.float32(c['bn']['one_minus_decay'])
        if c['residual_from'] is None:
            assert l['residual'] == -1
        else:
            src = L[l['residual']]['name']
            assert c['residual_from'] in (src, src.rsplit('/', 1)[0] + '/add')


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import ctypes as C
    cfg = nat.AmsConfig()
    cfg.num_classes, cfg.height, cfg.width, cfg.class_count = 19, 64, 128, 1
    assert not nat.lib().ams_create(C.byref(cfg))
    assert 'no CPU fallback' in nat.last_error()
    from ams_b200.student import Student
    with pytest.raises(nat.NativeError):
        Student(19, 64, 128, [0])


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'ams_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'student_oracle' not in src and 'metagraph_interp' not in src and 'import oracle' not in src, f


def _dw_geometries():
    """(H, W, C, Ho, Wo, stride, dilation) of the 17 depthwise layers at several frame sizes (TF SAME geometry)."""
    blocks = [(32, 1, 1), (96, 2, 1), (144, 1, 1), (144, 2, 1), (192, 1, 1), (192, 1, 1), (192, 2, 1), (384, 1, 1), (384, 1, 1),
              (384, 1, 1), (384, 1, 1), (576, 1, 1), (576, 1, 1), (576, 1, 1), (960, 1, 2), (960, 1, 2), (960, 1, 2)]
    out = []
    for height in (64, 96, 256, 512):
        h, w = (height + 2) // 2, (2 * height + 2) // 2                 # stem output of the (H+1) x (2H+1) padded frame
        for c, s, d in blocks:
            ho, wo = -(-h // s), -(-w // s)
            out.append((h, w, c, ho, wo, s, d))
            h, w = ho, wo
    return out


@pytest.mark.parametrize('batch', [1, 8])
def test_depthwise_tile_planner_invariants(batch):
    """Host-only: for every depthwise geometry the planners pick a tile that covers the image, fits the shared-memory
    budget of its occupancy target, keeps stride-2 tile origins even, and the magic-number divisions the kernels use
    (n * (65536 / d + 1) >> 16 == n / d for n < 256) are exact for the chosen strip / tile widths."""
    import ctypes as C
    lib = nat.lib()
    out = (C.c_int * 8)()
    for (h, w, c, ho, wo, s, d) in _dw_geometries():
        assert lib.ams_debug_dw_tile(batch, h, w, c, ho, wo, s, d, out) == 0
        th, tw, ntx, nty, cb, smem, ctas, threads = list(out)
        assert c % cb == 0 and cb in (32, 48, 64) and threads in (252, 256)
        assert th * nty >= ho and tw * ntx >= wo and 0 < smem <= 72 * 1024
        assert ctas == batch * ntx * nty * (c // cb)
        nstrips = -(-tw // 4)
        iwp = (nstrips * 4 - 1) * s + 2 * d + 1
        for dd in (iwp, nstrips):
            assert all(((n * (65536 // dd + 1)) >> 16) == n // dd for n in range(256)), (dd,)
        assert lib.ams_debug_dw_bwd_tile(batch, h, w, c, ho, wo, s, d, out) == 0
        th, tw, ntx, nty, cb, smem, ctas, nstrips = list(out)
        assert th * nty >= h and tw * ntx >= w and 0 < smem <= 112 * 1024 and c % cb == 0      # two CTAs per SM, cp.async ring included
        if s == 2:
            assert th % 2 == 0 and tw % 2 == 0
        owp = nstrips * 4 + 2 * d if s == 1 else nstrips * 2 + 1
        for dd in (owp, nstrips):
            assert all(((n * (65536 // dd + 1)) >> 16) == n // dd for n in range(256)), (dd,)
