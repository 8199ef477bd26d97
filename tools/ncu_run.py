"""One profiled pass of the bench workload for ncu (run under `ncu --profile-from-start off ...`):
   python tools/ncu_run.py train   -> 3 distillation steps (batch 8 @ 512x1024, masked Adam), AMS_NO_GRAPH=1 so that every
                                      kernel is a separate launch for the launch list
   python tools/ncu_run.py infer   -> 3 frozen inference batches (batch 8)
   python tools/ncu_run.py teacher -> 1 teacher forward at 1025x2049
Warm-up happens before cudaProfilerStart."""
import os, sys
os.environ.setdefault('AMS_NO_GRAPH', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ams_b200 import _native as nat
mode = sys.argv[1] if len(sys.argv) > 1 else 'train'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if mode == 'teacher':
    from ams_b200.teacher import XceptionTeacher, synthetic_teacher_checkpoint
    t = XceptionTeacher(synthetic_teacher_checkpoint(19, 1), 19)
    fr = np.random.default_rng(0).integers(0, 256, size=(1, 1025, 2049, 3), dtype=np.uint8)
    t.predict_batch(fr)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    t.predict_batch(fr)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    t.close()
else:
    from ams_b200.student import Student
    from ams_b200.synthetic import synthetic_checkpoint, synthetic_frames, synthetic_labels
    H, W, B = 512, 1024, 8
    st = Student(19, H, W, [0, 1, 2, 8, 10, 11, 13], queue_capacity=4)
    for k, v in synthetic_checkpoint('cityscapes', 1).items():
        st.set_tensor(k, v)
    fr, lab = synthetic_frames(B, H, W, 0), synthetic_labels(B, H, W, 0)
    def one():
        st.enqueue(fr, lab)
        st.train_step(1e-3, True) if mode == 'train' else st.infer_metric(B, nat.BN_MOVING)
    for _ in range(3):
        one()
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    for _ in range(reps):
        one()
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    st.close()
