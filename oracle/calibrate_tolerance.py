"""TEST INFRASTRUCTURE.  Measures what bf16 activation/weight storage (fp32 accumulate) costs against
the fp32 and fp64 oracle, to set the tolerances written in tests/ (SURVEY 8c 'calibrate with fp64').
usage: python oracle/calibrate_tolerance.py [H W N]"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import student_oracle as so

def bf16(t): return t.to(torch.bfloat16).to(t.dtype)
def rw(c, w): return bf16(w) if (c['op'] == 'Conv2D' and c['kh'] == 1) else w

def main():
    H, W, N = (int(a) for a in (sys.argv[1:4] + ['256', '512', '1'][len(sys.argv) - 1:]))
    torch.set_num_threads(os.cpu_count())
    for tag in ('cityscapes',):
        spec = so.load_spec(tag)
        V = so.synthetic_variables(spec, 1)
        fr = so.synthetic_frames(N, H, W, 0).astype(np.float32)
        for mode in ('moving', 'batch'):
            p64 = {k: torch.tensor(v, dtype=torch.float64) for k, v in V.items()}
            p32 = {k: torch.tensor(v) for k, v in V.items()}
            s64, _ = so.forward(spec, p64, fr, bn_mode=mode, dtype=torch.float64)
            s32, _ = so.forward(spec, p32, fr, bn_mode=mode)
            sbf, _ = so.forward(spec, p32, fr, bn_mode=mode, round_act=bf16, round_weight=rw,
                                round_conv=bf16 if mode == 'batch' else None)
            f64 = so.full_res_logits(s64, H, W); f32 = so.full_res_logits(s32, H, W); fbf = so.full_res_logits(sbf, H, W)
            a64 = f64.argmax(3); 
            print(tag, mode, f'{N}x{H}x{W}', '|logit|max %.2f' % float(s64.abs().max()),
                  'fp32-vs-fp64 maxabs %.2e' % float((s32.double() - s64).abs().max()),
                  'bf16-vs-fp64 maxabs %.3e mean %.3e' % (float((sbf.double() - s64).abs().max()), float((sbf.double() - s64).abs().mean())),
                  'argmax agree fp32 %.5f bf16 %.5f' % (float((f32.argmax(3) == a64).float().mean()), float((fbf.argmax(3) == a64).float().mean())))
if __name__ == '__main__':
    main()
