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
