"""No-GPU checks of the teacher path (config C5): the variable table of the restated DeepLabv3+ / Xception-65 graph in
the C library (ams_b200/csrc/teacher.cu, host-only layout query) against the oracle's INDEPENDENT restatement of the
same public model-zoo definition (oracle/teacher_oracle.py), and the oracle's own shape arithmetic."""
import numpy as np
import torch

import teacher_oracle as to
from ams_b200.teacher import synthetic_teacher_checkpoint, teacher_variables


def test_variable_tables_of_the_two_restatements_agree():
    for nc in (19, 21):
        lib = teacher_variables(nc)
        ora = to.variables(nc)
        assert len(lib) == len(ora) == 732
        assert sorted(lib) == sorted(ora)                     # same names and shapes (creation order differs: shortcut first in TF)
        assert sum(int(np.prod(s)) for _, s in lib) == (41257699 if nc == 19 else 41257699 + 2 * 257)
    # names a checkpoint of the public definition carries
    names = {n for n, _ in teacher_variables(19)}
    for n in ('xception_65/entry_flow/conv1_1/weights:0', 'xception_65/entry_flow/block2/unit_1/xception_module/separable_conv2_pointwise/weights:0',
              'xception_65/middle_flow/block1/unit_16/xception_module/separable_conv3_depthwise/depthwise_weights:0',
              'xception_65/exit_flow/block2/unit_1/xception_module/separable_conv3_pointwise/BatchNorm/moving_variance:0',
              'aspp3_depthwise/depthwise_weights:0', 'image_pooling/weights:0', 'concat_projection/weights:0',
              'decoder/feature_projection0/weights:0', 'decoder/decoder_conv1_pointwise/weights:0', 'logits/semantic/biases:0'):
        assert n in names, n


def test_synthetic_checkpoints_match_between_product_and_oracle():
    a = synthetic_teacher_checkpoint(19, seed=4)
    assert set(a) == set(n for n, _ in to.variables(19))
    assert a['concat_projection/weights:0'].shape == (1, 1, 1280, 256) and a['decoder/decoder_conv0_depthwise/depthwise_weights:0'].shape == (3, 3, 304, 1)


def test_oracle_shapes_at_the_teacher_strides():
    """33 x 49 input: output stride 16 features 3 x 4, decoder / logits at output stride 4 (9 x 13), labels at full size."""
    V = to.synthetic_checkpoint(19, seed=2)
    fr = np.random.default_rng(0).integers(0, 256, size=(1, 33, 49, 3), dtype=np.uint8)
    with torch.no_grad():
        logits, pred = to.forward(V, fr)
    assert tuple(logits.shape) == (1, 9, 13, 19) and tuple(pred.shape) == (1, 33, 49)
    assert torch.isfinite(logits).all() and int(pred.max()) < 19
