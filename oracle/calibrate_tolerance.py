"""TEST INFRASTRUCTURE.  Measures what the CUDA path's declared storage precision (bf16 activations and 1x1
weights, fp32 accumulate; oracle precision='bf16') costs against the fp32 and fp64 oracle, to set and justify
the tolerances written in tests/ ('calibrate with the fp64 oracle').
usage: python oracle/calibrate_tolerance.py [H W N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import student_oracle as so  # noqa: E402


def main():
    H, W, N = (int(a) for a in (sys.argv[1:4] + ['128', '256', '4'][len(sys.argv) - 1:]))
    torch.set_num_threads(os.cpu_count())
    spec = so.load_spec('cityscapes')
    fr = so.synthetic_frames(N, H, W, 0).astype(np.float32)
    for cond in (True, False):
        V = so.calibrate_moving_stats(spec, so.synthetic_variables(spec, 1, conditioned=cond), fr)
        p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in V.items()}
        p32 = {k: torch.tensor(v) for k, v in V.items()}
        for mode in ('moving', 'batch'):
            with torch.no_grad():
                s64, _ = so.forward(spec, p64, fr, bn_mode=mode, dtype=torch.float64)
                s32, _ = so.forward(spec, p32, fr, bn_mode=mode)
                sbf, _ = so.forward(spec, p32, fr, bn_mode=mode, precision='bf16')
            a64 = so.full_res_logits(s64, H, W).argmax(3)
            rel = lambda a: float((a.double() - s64).norm() / s64.norm())
            mx = lambda a: float((a.double() - s64).abs().max())
            agree = lambda a: float((so.full_res_logits(a, H, W).argmax(3) == a64).float().mean())
            print('conditioned=%d %-6s %dx%dx%d |logit| rms %.2f | fp32 vs fp64: rel %.1e max %.1e agree %.5f | '
                  'bf16 vs fp64: rel %.4f max %.3f agree %.4f' % (cond, mode, N, H, W, float(s64.pow(2).mean().sqrt()),
                                                                  rel(s32), mx(s32), agree(s32), rel(sbf), mx(sbf), agree(sbf)))


if __name__ == '__main__':
    main()
