"""BUILD INFRASTRUCTURE -- dump the reference's per-video experiment tables (pure data: class masks, video
lengths, label-space sizes, COCO->VOC map) to ams_b200/data/exp_configs.json by IMPORTING
/root/reference/exp_configs.py (reference `exp_configs.py:8-339`).  Runs only in the build container."""
import importlib.util
import json
import os

spec = importlib.util.spec_from_file_location('ref_exp_configs', '/root/reference/exp_configs.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
out = {'class_weights': {}, 'test_length': {}, 'num_classes': {}, 'is_coco': [],
       'coco_class_converter': ref.coco_class_converter().tolist()}
for e in range(0, 100):
    for key, fn in (('class_weights', ref.class_weights), ('test_length', ref.test_length), ('num_classes', ref.num_classes)):
        try:
            v = fn(e)
        except Exception:
            continue
        if v is None:
            continue
        out[key][str(e)] = v.reshape(-1).astype(int).tolist() if hasattr(v, 'reshape') else int(v)
    if ref.is_coco(e):
        out['is_coco'].append(e)
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ams_b200', 'data')
os.makedirs(dst, exist_ok=True)
with open(os.path.join(dst, 'exp_configs.json'), 'w') as f:
    json.dump(out, f)
print({k: (len(v) if hasattr(v, '__len__') else v) for k, v in out.items()})
