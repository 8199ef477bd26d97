"""Per-layer device timing of one distillation step / one frozen inference batch of the bench workload.
Writes a table (kernel group x layer: launches, us per launch, algorithmic GB/s) for profiles/."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ams_b200 import _native as nat
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels
out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/layers.txt'
H, W, B = 512, 1024, 8
st = Student(19, H, W, [0, 1, 2, 8, 10, 11, 13], queue_capacity=4)
for k, v in synthetic_checkpoint('cityscapes', 1).items():
    st.set_tensor(k, v)
fr, lab = synthetic_frames(B, H, W, 0), synthetic_labels(B, H, W, 0)
lines = []
for mode in ('train', 'infer'):
    for i in range(3):
        st.enqueue(fr, lab)
        st.train_step(1e-3, True) if mode == 'train' else st.infer_metric(B, nat.BN_MOVING)
    st.synchronize()
    st.profile_enable(2)
    reps = 5
    for i in range(reps):
        st.enqueue(fr, lab)
        st.train_step(1e-3, True) if mode == 'train' else st.infer_metric(B, nat.BN_MOVING)
    st.synchronize()
    rep = st.profile_report()
    st.profile_enable(0)
    tot = sum(v['ms'] for v in rep.values()) / reps
    lines.append('== %s: %.3f ms per step/batch (sum of kernel groups)' % (mode, tot))
    lines.append('%-16s %-42s %4s %9s %9s %8s' % ('group', 'layer', 'n', 'us', 'MB', 'GB/s'))
    for tag, v in rep.items():
        g, _, layer = tag.partition('@')
        us = v['ms'] * 1e3 / reps
        mb = v['algo_bytes'] / reps / 1e6
        lines.append('%-16s %-42s %4d %9.1f %9.2f %8.0f' % (g, layer, v['launches'] // reps, us, mb, mb / us * 1e3 if us else 0))
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:3]))
