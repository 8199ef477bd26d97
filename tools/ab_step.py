"""A/B helper: graph-replayed distillation step time (batch 8 @ 512x1024) + per-kernel-group times under the current
environment knobs (AMS_*).  Prints one JSON line; run it once per knob setting on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ams_b200.student import Student
from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels
H, W, B, K = 512, 1024, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = Student(19, H, W, [0, 1, 2, 8, 10, 11, 13], queue_capacity=K + 2)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st.set_stream(stream.cuda_stream)
for k, v in synthetic_checkpoint('cityscapes', 1).items():
    st.set_tensor(k, v)
batches = [(synthetic_frames(B, H, W, i), synthetic_labels(B, H, W, i)) for i in range(3)]
slots = torch.zeros(K + 8, dtype=torch.float32).pin_memory().numpy()
for i in range(4):
    st.enqueue(*batches[i % 3]); st.train_step_async(1e-3, True, slots[i:i + 1])
st.synchronize()
best = 1e9
for rep in range(2):
    for i in range(K):
        st.enqueue(*batches[i % 3])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for i in range(K):
        st.train_step_async(1e-3, True, slots[i:i + 1])
    e1.record(stream); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / K)
st.profile_enable(True)
for i in range(3):
    st.enqueue(*batches[i % 3]); st.train_step(1e-3, True)
prof = st.profile_report(); st.profile_enable(False)
knobs = {k: v for k, v in os.environ.items() if k.startswith('AMS_')}
print(json.dumps({'knobs': knobs, 'ms_per_step': round(best, 4), 'loss': float(slots[K - 1]),
                  'groups_ms': {k: round(v['ms'] / 3, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:12]}}))
st.close()
