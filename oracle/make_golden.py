"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by EXECUTING the reference's shipped model.meta with
oracle/metagraph_interp.py (runs only where /root/reference exists; the fixtures travel to the GPU box).

Each fixture holds: a seeded synthetic checkpoint digest (regenerated at test time from the same seed), the input
frames, and the interpreter's outputs (`semantic`, `student_logits` samples, moving-average updates, per-layer
activation checksums) for both BatchNorm modes.  tests/test_oracle.py re-computes them with the hand-written
restatement (oracle/student_oracle.py) and requires bit-equality in fp32.

usage: python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import student_oracle as so  # noqa: E402
from metagraph_interp import MetaGraphInterpreter  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
REF = '/root/reference/checkpoints'
CASES = [('cityscapes', 'deeplabv3_mobilenetv2_cityscapes', 2, 48, 80, True),
         ('cityscapes', 'deeplabv3_mobilenetv2_cityscapes', 1, 64, 128, False),
         ('pascalvoc2012', 'deeplabv3_mobilenetv2_pascalvoc2012', 2, 33, 47, True)]


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for tag, d, n, h, w, cond in CASES:
        spec = so.load_spec(tag)
        it = MetaGraphInterpreter(os.path.join(REF, d, 'model.meta'))
        V = so.synthetic_variables(spec, seed=7, conditioned=cond)
        frames = so.synthetic_frames(n, h, w, seed=11)
        fr = frames.astype(np.float32)
        out = {'frames': frames, 'seed': np.array(7), 'conditioned': np.array(int(cond))}
        for mode, ov in (('moving', 'moving'), ('batch', None)):
            fetch = ['semantic', 'student_logits', 'MobilenetV2/expanded_conv_3/project/BatchNorm/FusedBatchNormV3',
                     'MobilenetV2/expanded_conv_14/depthwise/Relu6', 'concat_projection/Relu']
            sem, full, a3, a14, cp = it.run(fetch, V, features=fr, bn_override=ov)
            out[mode + '/semantic'] = sem.numpy()
            out[mode + '/student_logits_argmax'] = full.argmax(3).numpy().astype(np.uint8)
            out[mode + '/student_logits_row0'] = full[:, 0].numpy()
            out[mode + '/expanded_conv_3_project_bn'] = a3.numpy()
            out[mode + '/expanded_conv_14_depthwise_relu6_sum'] = np.array(a14.double().sum().item())
            out[mode + '/concat_projection_relu_sum'] = np.array(cp.double().sum().item())
        upd = it.moving_average_updates(V, fr)
        for k in ('MobilenetV2/Conv/BatchNorm/moving_mean:0', 'MobilenetV2/expanded_conv_7/depthwise/BatchNorm/moving_variance:0',
                  'aspp0/BatchNorm/moving_variance:0', 'image_pooling/BatchNorm/moving_mean:0'):
            out['update/' + k] = upd[k]
        p = os.path.join(OUT, 'interp_%s_%dx%dx%d.npz' % (tag, n, h, w))
        np.savez_compressed(p, **out)
        print(p, os.path.getsize(p) // 1024, 'KB')


if __name__ == '__main__':
    main()
