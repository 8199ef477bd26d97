"""Runs W warm-up + a few distillation steps / inference batches of the bench workload (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ams_b200 import _native as nat
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels
mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W, B = 512, 1024, 8
st = Student(19, H, W, [0, 1, 2, 8, 10, 11, 13], queue_capacity=4)
for k, v in synthetic_checkpoint('cityscapes', 1).items():
    st.set_tensor(k, v)
fr, lab = synthetic_frames(B, H, W, 0), synthetic_labels(B, H, W, 0)
for i in range(steps):
    st.enqueue(fr, lab)
    if mode == 'train':
        st.train_step(1e-3, True)
    else:
        st.infer_metric(B, nat.BN_MOVING)
st.synchronize()
print('done', mode, steps)
