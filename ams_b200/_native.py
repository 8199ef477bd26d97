"""ctypes binding of libams_b200.so (the C ABI declared in include/ams_b200.h).

There is NO fallback: if the shared library is missing, was not built for this machine, or no sm_100a
device is present, importing/creating fails loudly.  Nothing here touches the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libams_b200.so')

AMS_MAX_CLASSES = 32
BN_MOVING, BN_BATCH = 0, 1
FRAMES_U8, FRAMES_F32 = 0, 1


class AmsConfig(C.Structure):
    _fields_ = [('num_classes', C.c_int), ('graph_variant', C.c_int), ('height', C.c_int), ('width', C.c_int),
                ('device', C.c_int), ('class_count', C.c_int), ('class_indices', C.c_int * AMS_MAX_CLASSES),
                ('label_depth', C.c_int), ('queue_capacity', C.c_int)]


class NativeError(RuntimeError):
    pass


_lib = None

_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_SIGNATURES = {
    'ams_last_error': (C.c_char_p, []),
    'ams_abi_version': (_i, []),
    'ams_launch_count': (_ll, []),
    'ams_create': (_vp, [C.POINTER(AmsConfig)]),
    'ams_destroy': (None, [_vp]),
    'ams_set_stream': (_i, [_vp, _vp]),
    'ams_synchronize': (_i, [_vp]),
    'ams_num_tensors': (_i, [_vp]),
    'ams_tensor_info': (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_ll)]),
    'ams_set_tensor': (_i, [_vp, C.c_char_p, _vp, _ll]),
    'ams_get_tensor': (_i, [_vp, C.c_char_p, _vp, _ll]),
    'ams_trainable_count': (_ll, [_vp]),
    'ams_get_trainable': (_i, [_vp, _vp]),
    'ams_set_trainable': (_i, [_vp, _vp]),
    'ams_reset_optimizer': (_i, [_vp]),
    'ams_enqueue': (_i, [_vp, _vp, _i, _vp, _i]),
    'ams_enqueue_raw': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i]),
    'ams_export_frozen': (_i, [_vp, C.c_char_p]),
    'ams_create_frozen': (_vp, [C.c_char_p, _vp]),
    'ams_is_frozen': (_i, [_vp]),
    'ams_set_block_fusion': (_i, [_vp, _i]),
    'ams_set_infer_split': (_i, [_vp, _i]),
    'ams_queue_size': (_i, [_vp]),
    'ams_queue_clear': (_i, [_vp]),
    'ams_infer': (_i, [_vp, _i, _vp]),
    'ams_infer_metric': (_i, [_vp, _i, _vp, _vp, C.POINTER(_f)]),
    'ams_confmat_labels': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'ams_train_step': (_i, [_vp, _f, _i, C.POINTER(_f)]),
    'ams_set_mask': (_i, [_vp, _vp]),
    'ams_get_mask': (_i, [_vp, _vp]),
    'ams_snapshot_before': (_i, [_vp]),
    'ams_select_topk': (_i, [_vp, _d, C.POINTER(_ll), C.POINTER(_f)]),
    'ams_pack_delta': (_i, [_vp, _vp, _ll, C.POINTER(_ll)]),
    'ams_train_forward_backward': (_i, [_vp, C.POINTER(_ll), C.POINTER(_d)]),
    'ams_gradient_arena': (_vp, [_vp, C.POINTER(_ll)]),
    'ams_apply_optimizer': (_i, [_vp, _f, _i, _f]),
    'ams_gradient_bucket_split': (_ll, [_vp]),
    'ams_gradient_bucket_wait': (_i, [_vp, _vp]),
    'ams_train_step_async': (_i, [_vp, _f, _i, _vp]),
    'ams_step_terms_device': (_vp, [_vp]),
    'ams_apply_optimizer_device': (_i, [_vp, _f, _i, _vp]),
    'ams_syncbn_init': (_i, [_vp, _i, _i, _vp, _i]),
    'ams_syncbn_connect': (_i, [_vp, _vp, _i]),
    'ams_syncbn_enable': (_i, [_vp, _i]),
    'ams_syncbn_status': (_i, [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    'ams_apply_delta': (_i, [_vp, _vp, _ll, C.POINTER(_ll)]),
    'ams_teacher_create': (_vp, [_i, _i]),
    'ams_teacher_destroy': (None, [_vp]),
    'ams_teacher_num_tensors': (_i, [_vp]),
    'ams_teacher_tensor_info': (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i)]),
    'ams_teacher_set_tensor': (_i, [_vp, C.c_char_p, _vp, _ll]),
    'ams_teacher_predict': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    'ams_teacher_time_forward': (_i, [_vp, _i, C.POINTER(_f)]),
    'ams_teacher_layout_num_tensors': (_i, [_i]),
    'ams_teacher_layout_tensor_info': (_i, [_i, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i)]),
    'ams_get_logits': (_i, [_vp, _vp, _ll]),
    'ams_get_gradients': (_i, [_vp, _vp]),
    'ams_num_layers': (_i, [_vp]),
    'ams_layer_info': (_i, [_vp, _i, C.c_char_p, _i] + [C.POINTER(_i)] * 6 + [C.POINTER(_f)] * 2 + [C.POINTER(_i)]),
    'ams_get_activation': (_i, [_vp, _i, _i, _vp, _ll]),
    'ams_profile_enable': (_i, [_vp, _i]),
    'ams_profile_report': (_i, [_vp, C.c_char_p, _i]),
    'ams_layout_num_tensors': (_i, [_i, _i]),
    'ams_layout_tensor_info': (_i, [_i, _i, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_ll)]),
    'ams_layout_num_layers': (_i, [_i, _i]),
    'ams_layout_layer_info': (_i, [_i, _i, _i, C.c_char_p, _i] + [C.POINTER(_i)] * 6 + [C.POINTER(_f)] * 2 + [C.POINTER(_i)]),
    'ams_debug_dw_tile': (_i, [_i] * 8 + [C.POINTER(_i)]),
    'ams_debug_dw_bwd_tile': (_i, [_i] * 8 + [C.POINTER(_i)]),
    'ams_op_conv1x1': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    'ams_op_fused_block': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'ams_debug_fused_timeline': (_i, [_vp]),
    'ams_op_wgrad': (_i, [_vp, _i, _vp, _i, _ll, _vp, _vp]),
    'ams_op_depthwise': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    'ams_op_resize_u8': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    'ams_op_depthwise_fused': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    'ams_op_depthwise_bwd_fused': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'ams_op_depthwise_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'ams_op_stem': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'ams_op_stem_bwd': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'ams_op_bn_train': (_i, [_vp, _ll, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    'ams_op_bn_backward': (_i, [_vp, _vp, _ll, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    'ams_op_head_infer': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, C.POINTER(_d), C.POINTER(_ll), _vp]),
    'ams_op_head_backward': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, C.POINTER(_f), _vp]),
    'ams_op_select': (_i, [_vp, _vp, _ll, _d, _vp, C.POINTER(_ll), C.POINTER(_f), _vp]),
    'ams_op_adam': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _vp]),
}


def exported_symbols():
    """Every symbol include/ams_b200.h declares (used by the no-GPU ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    """Load libams_b200.so once; raise if it is not there (no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                              '(nvcc, sm_100a). ams_b200 has no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError if the ABI and the header disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().ams_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    if rc != 0:
        msg = last_error()
        if 'KeyError: ' in msg:                      # unknown variable name: same exception as utils/utils.py:41
            raise KeyError(msg.split('KeyError: ', 1)[1].split(' at ')[0])
        raise NativeError('%s failed: %s' % (what or 'libams_b200 call', msg))


def layout(num_classes, graph_variant):
    """Host-only: (variables, layers) tables of the restated graph -- no device needed."""
    L = lib()
    name = C.create_string_buffer(256)
    shape = (C.c_int * 4)()
    nd, tr, off = C.c_int(), C.c_int(), C.c_longlong()
    variables = []
    for i in range(L.ams_layout_num_tensors(num_classes, graph_variant)):
        check(L.ams_layout_tensor_info(num_classes, graph_variant, i, name, 256, shape, C.byref(nd), C.byref(tr), C.byref(off)))
        variables.append(dict(name=name.value.decode(), shape=list(shape[:nd.value]), trainable=bool(tr.value), offset=off.value))
    layers = []
    iv = [C.c_int() for _ in range(7)]
    fv = [C.c_float(), C.c_float()]
    for i in range(L.ams_layout_num_layers(num_classes, graph_variant)):
        check(L.ams_layout_layer_info(num_classes, graph_variant, i, name, 256, *[C.byref(v) for v in iv[:6]],
                                      C.byref(fv[0]), C.byref(fv[1]), C.byref(iv[6])))
        layers.append(dict(name=name.value.decode(), kind=iv[0].value, cin=iv[1].value, cout=iv[2].value, stride=iv[3].value,
                           dilation=iv[4].value, act=iv[5].value, eps=fv[0].value, one_minus_decay=fv[1].value,
                           residual=iv[6].value))
    return variables, layers
