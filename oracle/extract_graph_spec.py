"""TEST/BUILD INFRASTRUCTURE -- derive ams_b200/graphs/<name>.json from the reference's model.meta.

Runs only where /root/reference exists.  The JSON it writes is DERIVED DATA (layer table, variable
names/shapes/order, BN eps/decay, conv strides/dilations), not reference source; it is what
`utils/graph_utils.py:350` (tf.train.import_meta_graph) would hand to TensorFlow.  The product's
C++ topology builder (ams_b200/csrc/topology.cpp) is checked against this file by
tests/test_layout.py, and oracle/student_oracle.py builds its network from it.

usage: python oracle/extract_graph_spec.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from metagraph_interp import MetaGraphInterpreter  # noqa: E402

REF = '/root/reference/checkpoints'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ams_b200', 'graphs')


def extract(meta_path, tag):
    it = MetaGraphInterpreter(meta_path)
    nodes = it.nodes
    consumers = {}
    for n in it.mg.graph_def.node:
        for i in n.input:
            consumers.setdefault(i.lstrip('^').split(':')[0], []).append(n.name)

    def const(name):
        return it._const(nodes[name])

    def var_of(read_name):
        n = nodes[read_name]
        while n.op == 'Identity':
            n = nodes[n.input[0]]
        assert n.op == 'VariableV2', n.op
        return n.name + ':0'

    def producer_layer(name):
        """Walk back through Identity/Relu/BN/... to the conv (or AddV2 / input) that made this tensor."""
        n = nodes[name.split(':')[0]]
        while True:
            if n.op in ('Conv2D', 'DepthwiseConv2dNative'):
                return n.name.rsplit('/', 1)[0] if n.op == 'Conv2D' else n.name.rsplit('/', 1)[0]
            if n.op == 'AddV2' and n.name.endswith('/add'):
                return n.name
            if n.name == 'MobilenetV2/MobilenetV2/input':
                return 'input'
            if n.op in ('Mean', 'ConcatV2'):
                return n.name
            n = nodes[n.input[0].split(':')[0]]

    shapes = it.variable_shapes()
    convs = []
    for n in it.mg.graph_def.node:
        if n.op not in ('Conv2D', 'DepthwiseConv2dNative'):
            continue
        if n.name.startswith('gradients') or 'patch' in n.name:
            continue
        wname = var_of(n.input[1])
        kh, kw, cin, cout = shapes[wname]
        ent = {'name': n.name.rsplit('/', 1)[0], 'op': n.op, 'weights': wname,
               'kh': kh, 'kw': kw, 'cin': cin, 'cout': cout if n.op == 'Conv2D' else cin,
               'stride': int(n.attr['strides'].list.i[1]), 'dilation': int(n.attr['dilations'].list.i[1]),
               'padding': n.attr['padding'].s.decode()}
        src = nodes[n.input[0].split(':')[0]]
        out_name = n.name
        if src.op == 'SpaceToBatchND':
            block = const(src.input[1])
            assert block[0] == block[1]
            ent['dilation'] = int(block[0])
            ent['padding'] = 'SAME'  # atrous pattern SpaceToBatch -> VALID -> BatchToSpace == SAME dilated
            ent['atrous_via_space_to_batch'] = True
            (b2s,) = [c for c in consumers[n.name] if nodes[c].op == 'BatchToSpaceND']
            out_name = b2s
            src = nodes[src.input[0]]
        ent['input'] = producer_layer(src.name)
        # follow consumers: BN / BiasAdd / activation / residual
        cur = out_name
        ent['bn'] = None
        ent['bias'] = None
        ent['act'] = None
        ent['residual_from'] = None
        while True:
            cs = [c for c in consumers.get(cur, []) if not c.startswith('gradients')]
            nxt = None
            for c in cs:
                cn = nodes[c]
                if cn.op == 'FusedBatchNormV3' and 'patch' not in c:
                    base = c[:-len('FusedBatchNormV3')]
                    decay = None
                    for suffix in ('AssignMovingAvg/decay', 'Const_2'):
                        if base + suffix in nodes:
                            v = float(const(base + suffix))
                            decay = 1.0 - v if suffix.endswith('decay') else v
                            break
                    ent['bn'] = {'gamma': var_of(cn.input[1]), 'beta': var_of(cn.input[2]),
                                 'moving_mean': base + 'moving_mean:0',
                                 'moving_variance': base + 'moving_variance:0',
                                 'eps': float(np.float32(cn.attr['epsilon'].f)), 'decay': decay,
                                 'is_training': bool(cn.attr['is_training'].b)}
                    nxt = c
                elif cn.op == 'BiasAdd':
                    ent['bias'] = var_of(cn.input[1])
                    nxt = c
                elif cn.op in ('Relu6', 'Relu'):
                    ent['act'] = cn.op.lower()
                    nxt = c
                elif cn.op == 'Identity':
                    nxt = c
                elif cn.op == 'AddV2' and c.endswith('/add') and cn.input[0].split(':')[0] == cur:
                    other = [i for i in cn.input if i.split(':')[0] != cur][0]
                    ent['residual_from'] = producer_layer(other)
                    ent['residual_add_name'] = c
            if nxt is None:
                break
            cur = nxt
        convs.append(ent)

    # moving-average decay as actually wired: AssignSub(mv, (mv - batch) * k)  => decay = 1-k
    for ent in convs:
        if ent['bn'] is not None:
            base = ent['bn']['moving_mean'][:-len('moving_mean:0')]
            k = None
            for cand in (base + 'AssignMovingAvg/decay', base + 'AssignMovingAvg/sub'):
                if cand in nodes and nodes[cand].op == 'Const':
                    k = float(const(cand))
            if k is None and base + 'AssignMovingAvg/sub' in nodes:
                n = nodes[base + 'AssignMovingAvg/sub']
                a, b = [float(const(i)) for i in n.input]
                k = a - b
            ent['bn']['decay'] = float(np.float32(1.0 - k)) if k is not None else ent['bn']['decay']
            ent['bn']['one_minus_decay'] = k

    num_classes = shapes['logits/semantic/biases:0'][0]
    # input preprocessing constants (node names differ between the two shipped graphs: walk, don't name)
    def const_input(n):
        (c,) = [i for i in n.input if nodes[i].op == 'Const']
        (o,) = [i for i in n.input if nodes[i].op != 'Const']
        return float(const(c)), nodes[o]

    sub_n = nodes[nodes['MobilenetV2/MobilenetV2/input'].input[0]]
    assert sub_n.op == 'Sub'
    shift, mul_n = const_input(sub_n)
    assert mul_n.op == 'Mul'
    scale, pad_n = const_input(mul_n)
    pre = {'norm_scale': scale, 'norm_shift': shift}
    if pad_n.op == 'PadV2':
        pre['pad'] = const(pad_n.input[1]).tolist()
        pre['pad_value'] = float(const(pad_n.input[2]))
        pre['pad_impl'] = 'PadV2'
    else:
        assert pad_n.op == 'ConcatV2'
        pre['pad'] = [[0, 0], [0, 1], [0, 1], [0, 0]]
        fill_mul = nodes[pad_n.input[1]]
        v, fill = const_input(fill_mul) if fill_mul.op == 'Mul' else (None, None)
        pre['pad_value'] = v * float(const(fill.input[1]))
        pre['pad_impl'] = 'Fill+ConcatV2'
    # logits resize 1: size = int((in-1)*scale+1)
    r1 = nodes['ResizeBilinear_1']
    cast = nodes[nodes[r1.input[1]].input[0]]
    add = nodes[cast.input[0]]
    mul = nodes[[i for i in add.input if nodes[i].op == 'Mul'][0]]
    pre['logits_resize_scale'] = const_input(mul)[0]
    concat_n = nodes[nodes['concat_projection/Conv2D'].input[0]]
    assert concat_n.op == 'ConcatV2'
    spec = {
        'tag': tag,
        'tensorflow_version': it.mg.meta_info_def.tensorflow_version,
        'num_classes': int(num_classes),
        'preprocess': pre,
        'aspp_concat_inputs': [producer_layer(i) for i in concat_n.input[:-1]],
        'trainable_variables': [{'name': v, 'shape': list(shapes[v])} for v in it.trainable],
        'variables': [{'name': v, 'shape': list(shapes[v])} for v in it.all_variables],
        'convs': convs,
    }
    return spec


def main():
    os.makedirs(OUT, exist_ok=True)
    for d, tag in (('deeplabv3_mobilenetv2_cityscapes', 'cityscapes'),
                   ('deeplabv3_mobilenetv2_pascalvoc2012', 'pascalvoc2012')):
        spec = extract(os.path.join(REF, d, 'model.meta'), tag)
        p = os.path.join(OUT, tag + '.json')
        with open(p, 'w') as f:
            json.dump(spec, f, indent=1)
        nt = sum(int(np.prod(v['shape'])) for v in spec['trainable_variables'])
        na = sum(int(np.prod(v['shape'])) for v in spec['variables'])
        print(tag, 'convs', len(spec['convs']), 'trainable', len(spec['trainable_variables']), nt,
              'all', len(spec['variables']), na, '->', p)


if __name__ == '__main__':
    main()
