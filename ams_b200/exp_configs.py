"""Per-video experiment tables of the reference (`exp_configs.py:8-339`): which classes a video is scored on,
its length, its label space, the COCO->PASCAL-VOC label map.  Pure data: stored in data/exp_configs.json
(dumped from the reference by tools/extract_exp_configs.py), served through the reference's function names,
argument meaning and error behaviour (ValueError for an unconfigured experiment)."""
import json
import os

import numpy as np

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'exp_configs.json')) as _f:
    _T = json.load(_f)


def _lookup(table, experiment_number):
    try:
        return _T[table][str(int(experiment_number))]
    except KeyError:
        raise ValueError('Experiment %d not configured' % experiment_number)


def num_classes(experiment_number):
    return _lookup('num_classes', experiment_number)


def class_weights(experiment_number):
    """0/1 column vector [num_classes, 1] float32 (reference exp_configs.py:18-199)."""
    v = np.array(_lookup('class_weights', experiment_number), dtype=np.float32)
    return np.reshape(v, (num_classes(experiment_number), 1))


def test_length(experiment_number):
    return _lookup('test_length', experiment_number)


def coco_class_converter():
    return np.array(_T['coco_class_converter'], dtype=np.int32)


def is_coco(experiment_number):
    return experiment_number in _T['is_coco']


def all_classes(n=19):
    """Not in the reference (its experiment 0 is unreachable: num_classes(0) raises): every class selected."""
    return np.ones((n, 1), dtype=np.float32)
