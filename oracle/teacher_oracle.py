"""TEST INFRASTRUCTURE ONLY.  CPU oracle (torch) of the teacher network of the label-extraction path (config C5).

PARITY STATUS: **parity unpinned AND unsourced.**  The reference runs the teacher through TensorFlow 1.15 on a graph it
imports from `<teacher_checkpoint>.meta` (utils/graph_utils.py:129-152, extract_labels.py:51-84); neither that .meta
nor the weights are in /root/reference (external download, README.md:45-46), and the reference holds no golden
vectors.  What this file restates is the PUBLIC model-zoo definition the checkpoint was exported from
(tensorflow/models, research/deeplab: core/xception.py `xception_65` + `xception_module` + `separable_conv2d_same`,
model.py `extract_features` (ASPP with image-level feature, atrous rates 6/12/18 at output stride 16) +
`refine_by_decoder` (decoder output stride 4) + `get_branch_logits`), independently of ams_b200/csrc/teacher.cu, so that
the two restatements check each other: variable names / shapes (tests/test_teacher.py, no GPU) and logits / labels
(tests/test_teacher_gpu.py).

Only tests/ may import this module."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

import student_oracle as so

BACKBONE_EPS, HEAD_EPS = 1e-3, 1e-5          # xception_arg_scope batch_norm_epsilon / model.py batch_norm_params


def _bn_vars(scope, c):
    return [(scope + '/BatchNorm/' + k + ':0', (c,)) for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')]


def variables(num_classes=19):
    """[(name, shape)] in graph-construction order of the public definition."""
    v = []
    x = 'xception_65/'
    v += [(x + 'entry_flow/conv1_1/weights:0', (3, 3, 3, 32))] + _bn_vars(x + 'entry_flow/conv1_1', 32)
    v += [(x + 'entry_flow/conv1_2/weights:0', (3, 3, 32, 64))] + _bn_vars(x + 'entry_flow/conv1_2', 64)

    def module(scope, cin, depth, skip):
        out = []
        if skip == 'conv':
            out += [(scope + '/shortcut/weights:0', (1, 1, cin, depth[2]))] + _bn_vars(scope + '/shortcut', depth[2])
        c = cin
        for i in range(3):
            s = '%s/separable_conv%d' % (scope, i + 1)
            out += [(s + '_depthwise/depthwise_weights:0', (3, 3, c, 1))] + _bn_vars(s + '_depthwise', c)
            out += [(s + '_pointwise/weights:0', (1, 1, c, depth[i]))] + _bn_vars(s + '_pointwise', depth[i])
            c = depth[i]
        return out
    for name, cin, depth, skip in blocks():
        v += module(name, cin, depth, skip)
    v += [('image_pooling/weights:0', (1, 1, 2048, 256))] + _bn_vars('image_pooling', 256)
    v += [('aspp0/weights:0', (1, 1, 2048, 256))] + _bn_vars('aspp0', 256)
    for i in (1, 2, 3):
        v += [('aspp%d_depthwise/depthwise_weights:0' % i, (3, 3, 2048, 1))] + _bn_vars('aspp%d_depthwise' % i, 2048)
        v += [('aspp%d_pointwise/weights:0' % i, (1, 1, 2048, 256))] + _bn_vars('aspp%d_pointwise' % i, 256)
    v += [('concat_projection/weights:0', (1, 1, 1280, 256))] + _bn_vars('concat_projection', 256)
    v += [('decoder/feature_projection0/weights:0', (1, 1, 256, 48))] + _bn_vars('decoder/feature_projection0', 48)
    for i, cin in ((0, 304), (1, 256)):
        s = 'decoder/decoder_conv%d' % i
        v += [(s + '_depthwise/depthwise_weights:0', (3, 3, cin, 1))] + _bn_vars(s + '_depthwise', cin)
        v += [(s + '_pointwise/weights:0', (1, 1, cin, 256))] + _bn_vars(s + '_pointwise', 256)
    v += [('logits/semantic/weights:0', (1, 1, 256, num_classes)), ('logits/semantic/biases:0', (num_classes,))]
    return v


def blocks():
    """(scope, input channels, depth_list, skip_connection_type) of every xception_module, in order (core/xception.py)."""
    x = 'xception_65/'
    out = [(x + 'entry_flow/block1/unit_1/xception_module', 64, (128, 128, 128), 'conv'),
           (x + 'entry_flow/block2/unit_1/xception_module', 128, (256, 256, 256), 'conv'),
           (x + 'entry_flow/block3/unit_1/xception_module', 256, (728, 728, 728), 'conv')]
    out += [(x + 'middle_flow/block1/unit_%d/xception_module' % u, 728, (728, 728, 728), 'sum') for u in range(1, 17)]
    out += [(x + 'exit_flow/block1/unit_1/xception_module', 728, (728, 1024, 1024), 'conv'),
            (x + 'exit_flow/block2/unit_1/xception_module', 1024, (1536, 1536, 2048), 'none')]
    return out


def synthetic_checkpoint(num_classes=19, seed=1):
    """Same recipe as ams_b200.teacher.synthetic_teacher_checkpoint, from THIS file's variable table."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in variables(num_classes):
        if name.endswith('weights:0'):
            fan_in = shape[0] * shape[1] * (1 if 'depthwise' in name else shape[2])
            a = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        elif name.endswith('gamma:0'):
            a = rng.uniform(0.5, 1.0, size=shape)
        elif name.endswith('beta:0'):
            a = rng.normal(0.3, 0.1, size=shape)
        elif name.endswith('moving_variance:0'):
            a = rng.uniform(0.5, 1.5, size=shape)
        else:
            a = rng.normal(0.0, 0.1, size=shape)
        out[name] = a.astype(np.float32)
    return out


def _fixed_pad_conv(x, w, stride, rate, depthwise):
    """conv2d_same / separable_conv2d_same (core/xception.py): stride 1 -> 'SAME'; stride > 1 -> explicit padding of
    k_eff - 1 (split beg = total // 2) followed by a 'VALID' convolution.  x NHWC, w HWIO."""
    k = w.shape[0]
    xn = x.permute(0, 3, 1, 2)
    k_eff = k + (k - 1) * (rate - 1)
    if stride == 1:
        p = (k_eff - 1) // 2
        xn = F.pad(xn, (p, k_eff - 1 - p, p, k_eff - 1 - p))
    else:
        beg = (k_eff - 1) // 2
        xn = F.pad(xn, (beg, k_eff - 1 - beg, beg, k_eff - 1 - beg))
    if depthwise:
        y = F.conv2d(xn, w.permute(2, 3, 0, 1), stride=stride, dilation=rate, groups=w.shape[2])
    else:
        y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=stride, dilation=rate)
    return y.permute(0, 2, 3, 1).contiguous()


def _bn(x, P, scope, eps):
    g, b = P[scope + '/BatchNorm/gamma:0'], P[scope + '/BatchNorm/beta:0']
    m, v = P[scope + '/BatchNorm/moving_mean:0'], P[scope + '/BatchNorm/moving_variance:0']
    return (x - m) * torch.rsqrt(v + eps) * g + b


def forward(variables_, frames, precision='fp32'):
    """frames [N,H,W,3] (0..255) -> (logits at output stride 4 [N,h,w,C], predictions int [N,H,W]).
    precision='fp16': the device's declared storage: every stored activation rounded to fp16, 1x1 weights
    fp16 (split hi + lo where the conv has at most 256 output channels), the 3x3 32->64 conv's weights fp16; fp32
    accumulation, BN folding, image-pooling branch and logits."""
    q = so._STORAGE_ROUND[precision]
    P = {k: torch.as_tensor(np.asarray(v)) for k, v in variables_.items()}
    x = torch.as_tensor(np.asarray(frames)).to(torch.float32)
    x = x * np.float32(2.0 / 255.0) - 1.0                                  # _preprocess_zero_mean_unit_range
    X = 'xception_65/'

    def pw(scope, t, act, eps, res=None, rows=None):
        w = P[scope + '/weights:0']
        if rows is not None:
            w = w[:, :, rows[0]:rows[1], :]
        w = so._weight_operand(w, precision) if precision != 'fp32' else w
        y = _bn(F.conv2d(t.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1)).permute(0, 2, 3, 1), P, scope, eps)
        if act:
            y = torch.relu(y)
        if res is not None:
            y = y + res
        return q(y)

    def dw(scope, t, stride, rate, pre_relu, act, eps):
        if pre_relu:
            t = torch.relu(t)
        y = _bn(_fixed_pad_conv(t, P[scope + '/depthwise_weights:0'], stride, rate, True), P, scope, eps)
        return q(torch.relu(y) if act else y)

    def sep(scope, t, stride, rate, act_inside, res=None):
        d = dw(scope + '_depthwise', t, stride, rate, not act_inside, act_inside, BACKBONE_EPS)
        return pw(scope + '_pointwise', d, act_inside, BACKBONE_EPS, res)

    # entry flow root: 3x3 s2 conv (fp32 weights on the device), then the dense 3x3 32 -> 64 (fp16 weights)
    x = q(torch.relu(_bn(_fixed_pad_conv(x, P[X + 'entry_flow/conv1_1/weights:0'], 2, 1, False), P, X + 'entry_flow/conv1_1', BACKBONE_EPS)))
    w12 = P[X + 'entry_flow/conv1_2/weights:0']
    if precision != 'fp32':
        w12 = q(w12)
    x = q(torch.relu(_bn(_fixed_pad_conv(x, w12, 1, 1, False), P, X + 'entry_flow/conv1_2', BACKBONE_EPS)))
    # stack_blocks_dense(output_stride=16): strides 2, 2, 2 (-> 16), then the stride of exit_flow/block1 turns into a rate
    plan = [2, 2, 2] + [1] * 16 + [(1, 1), (1, 2)]
    low = None
    for (scope, cin, depth, skip), st in zip(blocks(), plan):
        stride, rate = (st, 1) if not isinstance(st, tuple) else st
        act_inside = scope.endswith('exit_flow/block2/unit_1/xception_module')
        if skip == 'conv':
            src = x[:, ::2, ::2, :] if stride == 2 else x                # 1x1 'SAME' conv with stride 2
            shortcut = pw(scope + '/shortcut', src, False, BACKBONE_EPS)
        elif skip == 'sum':
            shortcut = x
        else:
            shortcut = None
        y = sep(scope + '/separable_conv1', x, 1, rate, act_inside)
        y = sep(scope + '/separable_conv2', y, 1, rate, act_inside)
        if scope.endswith('entry_flow/block2/unit_1/xception_module'):
            low = y                                                       # decoder end point: separable_conv2_pointwise
        x = sep(scope + '/separable_conv3', y, stride, rate, act_inside, shortcut)
    feat = x
    # ASPP
    pooled = feat.mean(dim=(1, 2), keepdim=True)
    ip = torch.relu(_bn(F.conv2d(pooled.permute(0, 3, 1, 2), P['image_pooling/weights:0'].permute(3, 2, 0, 1)).permute(0, 2, 3, 1),
                        P, 'image_pooling', HEAD_EPS))                    # [N,1,1,256], fp32 on the device
    branches = [pw('aspp0', feat, True, HEAD_EPS)]
    for i, r in ((1, 6), (2, 12), (3, 18)):
        d = dw('aspp%d_depthwise' % i, feat, 1, r, False, True, HEAD_EPS)
        branches.append(pw('aspp%d_pointwise' % i, d, True, HEAD_EPS))
    cat = torch.cat(branches, dim=3)
    wcp = P['concat_projection/weights:0']
    bias = F.conv2d(ip.permute(0, 3, 1, 2), wcp[:, :, :256, :].permute(3, 2, 0, 1)).permute(0, 2, 3, 1)     # resize of a 1x1 map = broadcast
    w_rest = wcp[:, :, 256:, :]
    w_rest = so._weight_operand(w_rest, precision) if precision != 'fp32' else w_rest
    y = F.conv2d(cat.permute(0, 3, 1, 2), w_rest.permute(3, 2, 0, 1)).permute(0, 2, 3, 1) + bias
    cp = q(torch.relu(_bn(y, P, 'concat_projection', HEAD_EPS)))
    # decoder
    up = q(so.resize_bilinear_align(cp, low.shape[1], low.shape[2]))
    lowp = pw('decoder/feature_projection0', low, True, HEAD_EPS)
    d = torch.cat([up, lowp], dim=3)
    for i in (0, 1):
        s = 'decoder/decoder_conv%d' % i
        d = dw(s + '_depthwise', d, 1, 1, False, True, HEAD_EPS)
        d = pw(s + '_pointwise', d, True, HEAD_EPS)
    wl = P['logits/semantic/weights:0']
    wl = so._weight_operand(wl, precision) if precision != 'fp32' else wl
    logits = F.conv2d(d.permute(0, 3, 1, 2), wl.permute(3, 2, 0, 1)).permute(0, 2, 3, 1) + P['logits/semantic/biases:0']
    full = so.resize_bilinear_align(logits, frames.shape[1], frames.shape[2])
    return logits, full.argmax(dim=3)
