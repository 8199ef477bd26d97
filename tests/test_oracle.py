"""CPU tests pinning the oracle (no GPU): (1) the hand-written restatement equals, bit for bit in fp32, what
oracle/metagraph_interp.py computed by executing the reference's shipped model.meta (tests/golden/*.npz, generated
by oracle/make_golden.py); (2) closed-form unit cases for every restated TF op (the reference has no tests or
golden vectors of its own: SURVEY.md section 4)."""
import glob
import os

import numpy as np
import pytest
import torch

import student_oracle as so

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'interp_*.npz')))


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_restatement_matches_model_meta_interpreter(path):
    g = np.load(path)
    tag = os.path.basename(path).split('_')[1]
    spec = so.load_spec(tag)
    V = so.synthetic_variables(spec, seed=int(g['seed']), conditioned=bool(int(g['conditioned'])))
    frames = g['frames']
    n, h, w, _ = frames.shape
    params = {k: torch.tensor(v) for k, v in V.items()}
    for mode in ('moving', 'batch'):
        keep = {}
        with torch.no_grad():
            sem, stats = so.forward(spec, params, frames.astype(np.float32), bn_mode=mode, keep=keep)
            full = so.full_res_logits(sem, h, w)
        assert np.array_equal(sem.numpy(), g[mode + '/semantic'])
        assert np.array_equal(full.argmax(3).numpy().astype(np.uint8), g[mode + '/student_logits_argmax'])
        assert np.array_equal(full[:, 0].numpy(), g[mode + '/student_logits_row0'])
        assert float(keep['MobilenetV2/expanded_conv_14/depthwise'].double().sum()) == float(g[mode + '/expanded_conv_14_depthwise_relu6_sum'])
        assert float(keep['concat_projection'].double().sum()) == float(g[mode + '/concat_projection_relu_sum'])
        assert np.array_equal(keep['MobilenetV2/expanded_conv_3/project'].numpy(), g[mode + '/expanded_conv_3_project_bn'])
    ts = so.TrainState(spec, V)
    ts.apply_moving_stats(stats)
    for k in g.files:
        if k.startswith('update/'):
            assert np.array_equal(ts.vars[k[len('update/'):]], g[k]), k


def test_bn_of_constant_image_is_beta():
    """FusedBatchNormV3(is_training): zero variance => y == beta; image_pooling at N == 1."""
    x = torch.full((1, 1, 1, 5), 3.25)
    gamma, beta = torch.rand(5) + 0.5, torch.randn(5)
    y, m, v = so.batch_norm(x, gamma, beta, 1e-3, 'batch')
    assert torch.allclose(y.reshape(-1), beta) and torch.equal(m, torch.full((5,), 3.25)) and float(v.abs().max()) == 0.0


def test_bn_unbiased_variance_output():
    x = torch.randn(2, 3, 4, 6)
    _, m, v = so.batch_norm(x, torch.ones(6), torch.zeros(6), 1e-3, 'batch')
    flat = x.reshape(-1, 6)
    assert torch.allclose(v, flat.var(dim=0, unbiased=True), atol=1e-6) and torch.allclose(m, flat.mean(0), atol=1e-6)


def test_same_padding_is_tf_style():
    # even input, stride 2, k=3: TF pads 0 before / 1 after (torch's symmetric padding would differ)
    assert so._same_pads(8, 3, 2, 1) == (0, 1)
    assert so._same_pads(513, 3, 2, 1) == (1, 1)
    assert so._same_pads(33, 3, 1, 2) == (2, 2)
    x = torch.arange(16, dtype=torch.float32).reshape(1, 4, 4, 1)
    w = torch.ones(3, 3, 1, 1)
    y = so.conv2d_same(x, w, 2, 1, False)
    assert y.shape == (1, 2, 2, 1) and float(y[0, 0, 0, 0]) == float(x[0, 0:3, 0:3, 0].sum())


def test_resize_align_corners_identity_and_corners():
    x = torch.randn(1, 5, 7, 3)
    assert torch.equal(so.resize_bilinear_align(x, 5, 7), x)
    y = so.resize_bilinear_align(x, 64, 128)
    for (a, b), (c, d) in (((0, 0), (0, 0)), ((63, 127), (4, 6)), ((0, 127), (0, 6)), ((63, 0), (4, 0))):
        assert torch.equal(y[0, a, b], x[0, c, d])
    lo, hi, l = so.resize_weights(33, 512)
    assert lo[0] == 0 and hi[-1] == 32 and lo[-1] == 32 and float(l[-1]) == 0.0 and np.all(hi - lo <= 1)


def test_confusion_matrix_and_miou_hand_case():
    labels = np.array([[0, 0, 1, 1, 2, 255]])
    pred = np.array([[0, 1, 1, 1, 0, 2]])
    fl, w = so.reduce_labels(labels, [0, 1, 2])
    cm = so.confusion_matrix(fl, pred, w, 3)
    assert np.array_equal(cm, np.array([[1, 1, 0], [0, 2, 0], [1, 0, 0]], dtype=np.float64))
    iou = so.calculate_miou(cm)
    assert iou[0] == 1 / 3 and iou[1] == 2 / 3 and iou[2] == 0.0
    # class neither present nor predicted => NaN, ignored by nanmean (SemanticNetwork.py:210-211)
    cm2 = so.confusion_matrix(*so.reduce_labels(np.array([[0, 0]]), [0, 5])[:1], np.array([[0, 0]]), np.ones((1, 2)), 2)
    assert np.isnan(so.calculate_miou(cm2)[1])


def test_reduce_labels_class_subset_and_out_of_range():
    cls = [0, 2, 13]
    fl, w = so.reduce_labels(np.array([0, 1, 2, 13, 18, 19, 200, 255]), cls)
    assert list(w) == [1, 0, 1, 1, 0, 0, 0, 0] and list(fl[:4]) == [0, 0, 1, 2]
    # one_hot depth is 19 even for the 21-class graph
    fl, w = so.reduce_labels(np.array([19, 20]), [19, 20])
    assert list(w) == [0, 0]


def test_loss_is_mean_over_valid_and_nan_when_empty():
    logits = torch.zeros(1, 2, 2, 19)
    h = so.head(logits, np.array([[[0, 255], [3, 3]]]), np.arange(19))
    assert h['n_valid'] == 3 and abs(float(h['loss']) - np.log(19)) < 1e-6
    h = so.head(logits, np.full((1, 2, 2), 255), np.arange(19))
    assert np.isnan(float(h['loss'])) and h['conf_mat'].sum() == 0


def test_adam_first_step_closed_form():
    """From fresh state TF1 Adam moves every coordinate by lr*g/(|g| + eps/sqrt(1-beta2))."""
    spec = {'trainable_variables': [{'name': 'v', 'shape': [5]}], 'convs': []}
    st = so.TrainState.__new__(so.TrainState)
    st.spec, st.trainable = spec, ['v']
    st.vars = {'v': np.zeros(5, np.float32)}
    st.m = {'v': np.zeros(5, np.float32)}
    st.v = {'v': np.zeros(5, np.float32)}
    st.beta1_power, st.beta2_power = so.BETA1, so.BETA2
    g = np.array([1e-3, -2.0, 5e-6, 0.0, 7.0], np.float32)
    st.adam_apply({'v': g}, 1e-3)
    expect = -1e-3 * g / (np.abs(g) + 1e-8 / np.sqrt(1 - 0.999))
    assert np.allclose(st.vars['v'], expect, rtol=2e-4, atol=1e-12)
    assert st.beta1_power == np.float32(0.9) * np.float32(0.9)


def test_masked_adam_keeps_unselected_but_updates_slots():
    spec = {'trainable_variables': [{'name': 'v', 'shape': [4]}], 'convs': []}
    st = so.TrainState.__new__(so.TrainState)
    st.spec, st.trainable = spec, ['v']
    st.vars = {'v': np.ones(4, np.float32)}
    st.m = {'v': np.zeros(4, np.float32)}
    st.v = {'v': np.zeros(4, np.float32)}
    st.beta1_power, st.beta2_power = so.BETA1, so.BETA2
    st.adam_apply({'v': np.ones(4, np.float32)}, 1e-3, {'v': np.array([True, False, True, False])})
    assert st.vars['v'][1] == 1.0 and st.vars['v'][0] < 1.0 and np.all(st.m['v'] > 0) and np.all(st.v['v'] > 0)


def test_percentile_selection_numpy119_semantics():
    rng = np.random.default_rng(0)
    a = rng.random(2113043).astype(np.float32)
    thr, lo, w_hi = so.percentile_threshold_np119(a, 0.05)
    assert lo == 2007389 and abs(w_hi - 0.9) < 1e-6                    # virtual index 2,007,389.9
    srt = np.sort(a)
    assert thr == np.float32(np.float64(srt[lo]) * (1 - w_hi) + np.float64(srt[lo + 1]) * w_hi)
    assert int((a > thr).sum()) == 105653                               # k = 105,653 kept when there are no ties
    # massive ties: strict '>' can select nothing
    t = np.full(1000, 1e-3, np.float32)
    thr, _, _ = so.percentile_threshold_np119(t, 0.1)
    assert int((t > thr).sum()) == 0


def test_pack_delta_wire_format():
    masks = [np.array([[1, 0, 0, 0, 0, 0, 0, 1], [1, 1, 0, 0, 0, 0, 0, 0]], bool), np.array([0, 1, 1], bool)]
    params = [np.arange(16, dtype=np.float32).reshape(2, 8), np.array([0.5, 1.5, 65536.0], np.float32)]
    blob = so.pack_delta(masks, params)
    assert blob[:3] == bytes([0b10000001, 0b11000000, 0b01100000])
    vals = np.frombuffer(blob[3:], dtype=np.float16)
    assert list(vals[:5]) == [0.0, 7.0, 8.0, 9.0, 1.5] and np.isinf(vals[5])


def test_frozen_graph_uses_patch_bn_epsilon_for_every_layer():
    """The frozen client graph (trim_graph_frozen(kill_norms=True), reference utils/graph_utils.py:362-369, :52-76)
    normalises with the `_patch` ops of tf.layers.batch_normalization(training=False): epsilon 1e-3 for ALL 54 layers,
    also image_pooling / aspp0 / concat_projection, whose training-graph epsilon is 1.001e-5."""
    spec = so.load_spec('cityscapes')
    aspp = [c for c in spec['convs'] if c['name'] in ('image_pooling', 'aspp0', 'concat_projection')]
    assert len(aspp) == 3 and all(abs(c['bn']['eps'] - 1.001e-5) < 1e-9 for c in aspp)
    x = torch.tensor([[[[2.0, -1.0]]]])
    gamma, beta = torch.tensor([1.5, 0.5]), torch.tensor([0.1, -0.2])
    mm, mv = torch.tensor([0.5, 0.25]), torch.tensor([1e-4, 4.0])          # a tiny moving variance makes epsilon visible
    y, _, _ = so.batch_norm(x, gamma, beta, 1.001e-5, 'moving', mm, mv)
    want = (x - mm) / torch.sqrt(mv + 1e-3) * gamma + beta
    assert torch.allclose(y, want, rtol=1e-6, atol=0)
    wrong = (x - mm) / torch.sqrt(mv + 1.001e-5) * gamma + beta
    assert abs(float(y[0, 0, 0, 0]) - float(wrong[0, 0, 0, 0])) > 1.0
    # batch mode keeps the node's own epsilon
    xb = torch.tensor([[[[1.0]], [[3.0]]]])
    yb, _, _ = so.batch_norm(xb, torch.ones(1), torch.zeros(1), 1.001e-5, 'batch')
    assert abs(float(yb[0, 1, 0, 0]) - 1.0 / np.sqrt(1.0 + 1.001e-5)) < 1e-6


def test_teacher_forcing_is_value_neutral():
    spec = so.load_spec('cityscapes')
    V = so.synthetic_variables(spec, 3)
    fr = so.synthetic_frames(1, 32, 48, 1).astype(np.float32)
    lab = so.synthetic_labels(1, 32, 48, 1, block=8)
    ts = so.TrainState(spec, V, precision=so.DEVICE_PRECISION)
    keep = {}
    with torch.no_grad():
        so.forward(spec, {k: torch.tensor(v) for k, v in V.items()}, fr, bn_mode='batch', precision=so.DEVICE_PRECISION, keep=keep)
    forced = {}
    for c in spec['convs']:
        if c['name'] in ('image_pooling', 'logits/semantic'):
            continue
        forced[c['name'] + '/z'] = keep[c['name'] + '/z']
        forced[c['name']] = keep[c['residual_add_name']] if c['residual_from'] else keep[c['name']]
    l1, g1, _, _ = ts.loss_and_grads(fr, lab, np.arange(19))
    l2, g2, _, _ = ts.loss_and_grads(fr, lab, np.arange(19), forced=forced)
    assert l1 == l2 and all(np.array_equal(g1[k], g2[k]) for k in g1)


# ----------------------------------------------------------------------------- frame ingest (cv2.resize) oracle
def test_resize_oracle_matches_cv2_golden_vectors():
    """The ingest oracle against outputs of OpenCV itself (tests/golden/resize_golden.npz, generated here by
    oracle/make_resize_golden.py): INTER_LINEAR uint8 and INTER_NEAREST, bit for bit."""
    import cv2_resize_oracle as ro
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'resize_golden.npz'))
    i = 0
    while 'src_%d' % i in g:
        dh, dw = g['linear_%d' % i].shape[:2]
        assert np.array_equal(ro.resize_linear_u8(g['src_%d' % i], dw, dh), g['linear_%d' % i]), i
        assert np.array_equal(ro.resize_nearest_u8(g['lab_%d' % i], dw, dh), g['nearest_%d' % i]), i
        i += 1
    assert i >= 7


def test_resize_oracle_matches_cv2_live_when_available():
    cv2 = pytest.importorskip('cv2')
    import cv2_resize_oracle as ro
    rng = np.random.default_rng(11)
    for (sh, sw, dh, dw) in [(270, 480, 128, 256), (120, 213, 128, 256), (256, 512, 128, 256), (97, 61, 40, 33), (30, 40, 64, 128)]:
        img = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
        lab = rng.integers(0, 19, size=(sh, sw), dtype=np.uint8)
        assert np.array_equal(ro.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh))), (sh, sw, dh, dw)
        assert np.array_equal(ro.resize_nearest_u8(lab, dw, dh), cv2.resize(lab, (dw, dh), interpolation=cv2.INTER_NEAREST))
        assert np.array_equal(ro.ingest_frame(img, dh, dw), cv2.cvtColor(cv2.resize(img, (dw, dh)), cv2.COLOR_BGR2RGB))


def test_syncbn_exchange_arithmetic_equals_single_process_batchnorm():
    """The cross-rank arithmetic of the data-parallel BN kernels (oracle/syncbn_oracle.py restates it) against torch
    autograd on the concatenated batch: statistics, Bessel-corrected variance, dz on every shard, d_gamma / d_beta after
    the gradient allreduce."""
    import syncbn_oracle as sb
    rng = np.random.default_rng(3)
    world, rows, C, eps = 4, 37, 24, 1e-3
    z = [rng.normal(0.3, 1.7, size=(rows, C)).astype(np.float32) for _ in range(world)]
    dy = [rng.normal(0.0, 1.0, size=(rows, C)).astype(np.float32) for _ in range(world)]
    gamma = rng.uniform(0.5, 1.5, size=C).astype(np.float32)
    beta = rng.normal(0, 0.1, size=C).astype(np.float32)
    mean, var, rstd, unb = sb.global_stats([sb.forward_sums(x) for x in z], rows, eps)
    zt = torch.tensor(np.concatenate(z), dtype=torch.float64, requires_grad=True)
    gt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    m = zt.mean(0)
    v = ((zt - m) ** 2).mean(0)
    y = (zt - m) / torch.sqrt(v + eps) * gt + bt
    y.backward(torch.tensor(np.concatenate(dy), dtype=torch.float64))
    n = rows * world
    assert np.allclose(mean, m.detach().numpy(), rtol=1e-6, atol=1e-7)
    assert np.allclose(var, v.detach().numpy(), rtol=1e-5) and np.allclose(unb, v.detach().numpy() * n / (n - 1), rtol=1e-5)
    pairs = [sb.backward_sums(g, x) for g, x in zip(dy, z)]
    dgs, dbs = np.zeros(C, np.float64), np.zeros(C, np.float64)
    for r in range(world):
        A, B, Cc, dg, db = sb.backward_coefficients(pairs, pairs[r], rows, mean, rstd, gamma)
        dz = A * dy[r] + B * z[r] + Cc
        assert np.allclose(dz, zt.grad.numpy()[r * rows:(r + 1) * rows], rtol=2e-4, atol=2e-5)
        dgs += dg
        dbs += db
    assert np.allclose(dgs, gt.grad.numpy(), rtol=2e-4, atol=1e-4) and np.allclose(dbs, bt.grad.numpy(), rtol=1e-5, atol=1e-5)
    # duplicated shards: all sums double and so does n -> statistics bit-identical to one shard alone (the exact GPU test)
    one = sb.global_stats([sb.forward_sums(z[0])], rows, eps)
    two = sb.global_stats([sb.forward_sums(z[0])] * 2, rows, eps)
    assert np.array_equal(one[0], two[0]) and np.array_equal(one[1], two[1]) and np.array_equal(one[2], two[2])
    # image-pooling BN: pooled (sum, centred sum of squares) == statistics over the concatenated batch dimension
    zp = [rng.normal(1.0, 2.0, size=(3, C)) for _ in range(world)]
    loc = [(x.sum(0), ((x - x.mean(0)) ** 2).sum(0)) for x in zp]
    pm, pv = sb.pooled_mean_var(loc, 3)
    allz = np.concatenate(zp)
    assert np.allclose(pm, allz.mean(0), rtol=1e-12) and np.allclose(pv, allz.var(0), rtol=1e-10)
