"""Frozen inference (batch 8 @ 512x1024, bench workload) timed with block fusion / batch split on and off, plus the
per-layer device times of the block-fused schedule.   usage: python tools/micro/infer_time.py [out.txt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from ams_b200 import _native as nat
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels
out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/infer_time.txt'
H, W, B = 512, 1024, 8
st = Student(19, H, W, [0, 1, 2, 8, 10, 11, 13], queue_capacity=4)
for k, v in synthetic_checkpoint('cityscapes', 1).items():
    st.set_tensor(k, v)
fr, lab = synthetic_frames(B, H, W, 0), synthetic_labels(B, H, W, 0)
lines = []
def timed(n=30):
    for _ in range(4):
        st.enqueue(fr, lab); st.infer_metric(B, nat.BN_MOVING)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        st.enqueue(fr, lab); st.infer_metric(B, nat.BN_MOVING)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for fusion in (False, True):
    for split in (False, True):
        st.set_block_fusion(fusion); st.set_infer_split(split)
        ms = timed()
        st.enqueue(fr, lab); pred, cm, loss = st.infer_metric(B, nat.BN_MOVING)
        if ref is None: ref = pred.copy()
        lines.append('block fusion %d, split %d: %.3f ms per batch of 8 incl. enqueue (H2D) and D2H  -> %.0f frames/s; argmax agreement with the per-layer schedule %.5f'
                     % (fusion, split, ms, B / ms * 1e3, float((pred == ref).mean())))
st.set_block_fusion(True); st.set_infer_split(False)
st.profile_enable(2)
reps = 5
for i in range(reps):
    st.enqueue(fr, lab); st.infer_metric(B, nat.BN_MOVING)
st.synchronize()
rep = st.profile_report()
st.profile_enable(0)
lines.append('== block-fused schedule: %.3f ms per batch (sum of kernel groups)' % (sum(v['ms'] for v in rep.values()) / reps))
for tag, v in rep.items():
    g, _, layer = tag.partition('@')
    us = v['ms'] * 1e3 / reps
    mb = v['algo_bytes'] / reps / 1e6
    lines.append('%-16s %-42s %4d %9.1f %9.2f %8.0f' % (g, layer, v['launches'] // reps, us, mb, mb / us * 1e3 if us else 0))
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
