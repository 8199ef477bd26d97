"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A small interpreter that EXECUTES the reference's shipped TF1 ``model.meta`` on the CPU
(torch fp32/fp64), node by node, with restated TensorFlow-1.15 op semantics.

It exists for exactly one purpose: to pin the hand-written oracle (``oracle/student_oracle.py``)
and the extracted graph spec (``ams_b200/graphs/*.json``) to the model definition the reference
actually ships (``/root/reference/checkpoints/*/model.meta``; imported by the reference at
``utils/graph_utils.py:350``).  It can only run where ``/root/reference`` exists (the build
container); ``oracle/make_golden.py`` uses it to generate ``tests/golden/*.npz``, which DO travel.

Parity status: the reference's arithmetic lives in TensorFlow 1.15.0 (``environment.yml:116-119``),
which is absent from ``/root/reference`` and not installable here, and the reference has no golden
vectors of its own => "parity unpinned" at the TF-kernel level.  What IS
pinned by this file: topology, constants, attrs (strides, dilations, eps, decay), variable
names/shapes/order -- all read from the reference's own protobuf, nothing hand-copied.

Op semantics restated (TF 1.15 kernels, public documentation):
  Conv2D / DepthwiseConv2dNative  NHWC, SAME => out=ceil(in/s), pad_total=max((out-1)s+(k-1)d+1-in,0),
                                  pad_before=pad_total//2 (extra at the end); VALID => no pad
  FusedBatchNormV3(is_training)   y=(x-mean)*rsqrt(var_biased+eps)*gamma+beta; outputs 1,2 = batch
                                  mean and UNBIASED batch variance (n/(n-1))
  ResizeBilinear(align_corners)   src=dst*(in-1)/(out-1); lo=floor(src); hi=min(ceil(src),in-1);
                                  lerp in x first, then y
  SpaceToBatchND / BatchToSpaceND block-shaped re-tiling with zero padding / cropping
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from tensorboard.compat.proto import meta_graph_pb2, variable_pb2
from tensorboard.util import tensor_util


def load_meta(path):
    mg = meta_graph_pb2.MetaGraphDef()
    with open(path, 'rb') as f:
        mg.ParseFromString(f.read())
    return mg


def collection_variable_names(mg, key):
    out = []
    for b in mg.collection_def[key].bytes_list.value:
        v = variable_pb2.VariableDef()
        v.ParseFromString(b)
        out.append(v.variable_name)
    return out


def _same_pads(in_size, k, s, d):
    out = -(-in_size // s)
    eff = (k - 1) * d + 1
    total = max((out - 1) * s + eff - in_size, 0)
    return total // 2, total - total // 2


def tf_conv2d(x, w, strides, padding, dilations, depthwise=False):
    """x NHWC, w HWIO (depthwise: [kh,kw,C,mult])."""
    sh, sw = int(strides[1]), int(strides[2])
    dh, dw_ = int(dilations[1]), int(dilations[2])
    kh, kw = w.shape[0], w.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == 'SAME':
        pt, pb = _same_pads(x.shape[1], kh, sh, dh)
        pl, pr = _same_pads(x.shape[2], kw, sw, dw_)
        xn = F.pad(xn, (pl, pr, pt, pb))
    if depthwise:
        c = w.shape[2]
        assert w.shape[3] == 1
        wt = w.permute(2, 3, 0, 1)  # [C,1,kh,kw]
        y = F.conv2d(xn, wt, stride=(sh, sw), dilation=(dh, dw_), groups=c)
    else:
        wt = w.permute(3, 2, 0, 1)
        y = F.conv2d(xn, wt, stride=(sh, sw), dilation=(dh, dw_))
    return y.permute(0, 2, 3, 1).contiguous()


def tf_resize_bilinear_align(x, out_h, out_w):
    """ResizeBilinear(align_corners=True), NHWC."""
    n, in_h, in_w, c = x.shape
    dt = x.dtype

    def weights(in_size, out_size):
        scale = (in_size - 1) / (out_size - 1) if out_size > 1 else 0.0
        scale = np.float32(scale)
        src = (np.arange(out_size, dtype=np.float32) * scale).astype(np.float32)
        lo = np.floor(src)
        hi = np.minimum(np.ceil(src), in_size - 1)
        lerp = (src - lo).astype(np.float32)
        return (torch.from_numpy(lo.astype(np.int64)), torch.from_numpy(hi.astype(np.int64)),
                torch.from_numpy(lerp).to(dt))

    ylo, yhi, yl = weights(in_h, out_h)
    xlo, xhi, xl = weights(in_w, out_w)
    top = x[:, ylo]       # [n,out_h,in_w,c]
    bot = x[:, yhi]
    xl_ = xl.view(1, 1, -1, 1)
    yl_ = yl.view(1, -1, 1, 1)
    top = top[:, :, xlo] + (top[:, :, xhi] - top[:, :, xlo]) * xl_
    bot = bot[:, :, xlo] + (bot[:, :, xhi] - bot[:, :, xlo]) * xl_
    return top + (bot - top) * yl_


def space_to_batch_nd(x, block, paddings):
    n, h, w, c = x.shape
    bh, bw = int(block[0]), int(block[1])
    (pt, pb), (pl, pr) = [[int(v) for v in r] for r in paddings]
    x = F.pad(x, (0, 0, pl, pr, pt, pb))
    hp, wp = x.shape[1], x.shape[2]
    x = x.reshape(n, hp // bh, bh, wp // bw, bw, c)
    x = x.permute(2, 4, 0, 1, 3, 5)
    return x.reshape(bh * bw * n, hp // bh, wp // bw, c).contiguous()


def batch_to_space_nd(x, block, crops):
    bh, bw = int(block[0]), int(block[1])
    nb, h, w, c = x.shape
    n = nb // (bh * bw)
    x = x.reshape(bh, bw, n, h, w, c).permute(2, 3, 0, 4, 1, 5)
    x = x.reshape(n, h * bh, w * bw, c)
    (ct, cb), (cl, cr) = [[int(v) for v in r] for r in crops]
    return x[:, ct:h * bh - cb, cl:w * bw - cr].contiguous()


class MetaGraphInterpreter:
    """Evaluate named tensors of the reference's graph.  ``variables`` maps '<name>:0' -> ndarray."""

    def __init__(self, meta_path, dtype=torch.float32):
        self.mg = load_meta(meta_path)
        self.nodes = {n.name: n for n in self.mg.graph_def.node}
        self.dtype = dtype
        self.trainable = collection_variable_names(self.mg, 'trainable_variables')
        self.all_variables = collection_variable_names(self.mg, 'variables')

    # -- helpers
    def _const(self, n):
        return tensor_util.make_ndarray(n.attr['value'].tensor)

    def variable_shapes(self):
        out = {}
        for name in self.all_variables:
            n = self.nodes[name[:-2]]
            out[name] = tuple(d.size for d in n.attr['shape'].shape.dim)
        return out

    def run(self, fetches, variables, features=None, labels=None, bn_override=None):
        """bn_override: None => execute FusedBatchNormV3 as the graph says (is_training=True);
        'moving' => frozen-client semantics of utils/graph_utils.py:52-76,362-369 (inference-mode BN
        bound to the original gamma/beta/moving_mean/moving_variance)."""
        cache = {}
        tt = lambda a: torch.as_tensor(np.asarray(a))

        def to_float(a):
            t = tt(a)
            return t.to(self.dtype) if t.dtype.is_floating_point else t

        def ev(ref):
            ref = ref.lstrip('^')
            name, _, idx = ref.partition(':')
            idx = int(idx) if idx else 0
            key = (name, idx)
            if key in cache:
                return cache[key]
            n = self.nodes[name]
            op = n.op
            I = lambda k: ev(n.input[k])
            if op == 'Const':
                v = self._const(n)
                out = (to_float(v) if v.dtype.kind == 'f' else v,)
            elif op == 'VariableV2':
                out = (to_float(variables[name + ':0']),)
            elif op in ('Identity', 'StopGradient', 'PlaceholderWithDefault'):
                out = (I(0),)
            elif op == 'QueueDequeueV2':
                # component 0 = labels, 1 = features (fill_input_buffer enqueues [labels_input, features_input])
                out = (to_float(labels) if labels is not None else None, to_float(features))
            elif op == 'Shape':
                out = (np.array(I(0).shape, dtype=np.int32),)
            elif op == 'StridedSlice':
                x, b, e, s = I(0), np.asarray(I(1)), np.asarray(I(2)), np.asarray(I(3))
                x = np.asarray(x)
                bm = n.attr['begin_mask'].i
                em = n.attr['end_mask'].i
                sm = n.attr['shrink_axis_mask'].i
                assert n.attr['ellipsis_mask'].i == 0 and n.attr['new_axis_mask'].i == 0
                sl = []
                for d in range(len(b)):
                    if sm & (1 << d):
                        sl.append(int(b[d]))
                    else:
                        bb = None if bm & (1 << d) else int(b[d])
                        ee = None if em & (1 << d) else int(e[d])
                        sl.append(slice(bb, ee, int(s[d])))
                out = (x[tuple(sl)],)
            elif op == 'Pack':
                vals = [np.asarray(ev(i)) for i in n.input]
                out = (np.stack(vals, axis=n.attr['axis'].i),)
            elif op == 'Fill':
                dims = [int(v) for v in np.asarray(I(0))]
                out = (torch.full(dims, float(np.asarray(I(1))), dtype=self.dtype),)
            elif op in ('Mul', 'Sub', 'Add', 'AddV2', 'FloorMod', 'Less'):
                a, b = I(0), I(1)
                if torch.is_tensor(a) or torch.is_tensor(b):
                    a = a if torch.is_tensor(a) else to_float(a)
                    b = b if torch.is_tensor(b) else to_float(b)
                f = {'Mul': lambda a, b: a * b, 'Sub': lambda a, b: a - b, 'Add': lambda a, b: a + b,
                     'AddV2': lambda a, b: a + b, 'FloorMod': lambda a, b: np.mod(a, b),
                     'Less': lambda a, b: a < b}[op]
                out = (f(a, b),)
            elif op == 'Cast':
                x = I(0)
                dst = n.attr['DstT'].type
                x = np.asarray(x)
                out = (x.astype(np.float32) if dst == 1 else x.astype(np.int32),)  # float->int truncates
            elif op == 'ConcatV2':
                axis = int(np.asarray(ev(n.input[-1])))
                parts = [ev(i) for i in n.input[:-1]]
                if all(not torch.is_tensor(p) for p in parts):
                    out = (np.concatenate([np.asarray(p) for p in parts], axis=axis),)
                else:
                    out = (torch.cat([p if torch.is_tensor(p) else to_float(p) for p in parts], dim=axis),)
            elif op == 'PadV2':
                x, pads, cv = I(0), np.asarray(I(1)), float(np.asarray(I(2)))
                flat = []
                for d in reversed(range(pads.shape[0])):
                    flat += [int(pads[d][0]), int(pads[d][1])]
                out = (F.pad(x, flat, value=cv),)
            elif op in ('Conv2D', 'DepthwiseConv2dNative'):
                out = (tf_conv2d(I(0), I(1), list(n.attr['strides'].list.i), n.attr['padding'].s.decode(),
                                 list(n.attr['dilations'].list.i), depthwise=(op != 'Conv2D')),)
            elif op == 'FusedBatchNormV3':
                x, gamma, beta = I(0), I(1), I(2)
                eps = n.attr['epsilon'].f
                if bn_override == 'moving':
                    # the frozen client graph swaps EVERY FusedBatchNormV3 for the `_patch` op that
                    # tf.layers.batch_normalization(training=False) creates (utils/graph_utils.py:362-369, :52-76):
                    # default epsilon 1e-3, also for the three ASPP layers whose training-graph epsilon is 1.001e-5
                    eps = 1e-3
                    base = name[:-len('FusedBatchNormV3')]
                    mean = to_float(variables[base + 'moving_mean:0'])
                    var = to_float(variables[base + 'moving_variance:0'])
                    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
                    out = (y, mean, var)
                else:
                    assert n.attr['is_training'].b
                    cnt = x.shape[0] * x.shape[1] * x.shape[2]
                    mean = x.mean(dim=(0, 1, 2))
                    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
                    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
                    out = (y, mean, var * (cnt / max(cnt - 1, 1)))
            elif op == 'Relu6':
                out = (torch.clamp(I(0), 0.0, 6.0),)
            elif op == 'Relu':
                out = (torch.clamp_min(I(0), 0.0),)
            elif op == 'Mean':
                axes = [int(a) for a in np.asarray(I(1)).reshape(-1)]
                out = (I(0).mean(dim=axes, keepdim=n.attr['keep_dims'].b),)
            elif op == 'BiasAdd':
                out = (I(0) + I(1),)
            elif op == 'ResizeBilinear':
                assert n.attr['align_corners'].b
                size = [int(v) for v in np.asarray(I(1))]
                out = (tf_resize_bilinear_align(I(0), size[0], size[1]),)
            elif op == 'SpaceToBatchND':
                out = (space_to_batch_nd(I(0), np.asarray(I(1)), np.asarray(I(2))),)
            elif op == 'BatchToSpaceND':
                out = (batch_to_space_nd(I(0), np.asarray(I(1)), np.asarray(I(2))),)
            else:
                raise NotImplementedError('%s (%s)' % (op, name))
            for i, v in enumerate(out):
                cache[(name, i)] = v
            return out[idx]

        import sys
        sys.setrecursionlimit(20000)
        return [ev(f) for f in fetches]

    def moving_average_updates(self, variables, features, labels=None):
        """Evaluate the 108 AssignSub nodes of collection 'update_ops': returns {var:0 -> new value}."""
        out = {}
        deltas = []
        names = []
        for n in self.mg.graph_def.node:
            if n.op == 'AssignSub':
                names.append(n.input[0] + ':0')
                deltas.append(n.input[1])
        vals = self.run(deltas, variables, features=features, labels=labels)
        for nm, d in zip(names, vals):
            out[nm] = (torch.as_tensor(variables[nm]).to(self.dtype) - d).numpy()
        return out
