"""GPU: the repo's own detectron2-configuration R-CNN graph (model/rcnn.py) and its glue kernels (csrc/rcnn.cu, nms.cu,
roi_align.cu) against the plain-PyTorch float32 restatement of detectron2's inference path (oracle/d2_rcnn_oracle.py) and
against the torchvision operators detectron2 itself calls.  Tolerances are written where they are asserted."""
import ctypes

import numpy as np
import pytest

import d2_rcnn_oracle as D

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
F = torch.nn.functional


@pytest.fixture(scope='module', autouse=True)
def _exact_float32():
    """float32 comparisons: no TF32 in cuDNN / cuBLAS for this module."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


@pytest.fixture(scope='module')
def state():
    return {k: v.cuda() for k, v in D.make_random_state(3).items()}


@pytest.fixture(scope='module')
def images():
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.proc import prep_raw_frames, scale_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(4, seed=9, geom=geom)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom),
                           vmin=0, vmax=100)
    scaled = scale_raw_frames(prep, 0, 100)
    return prep, [s[None].expand(3, -1, -1) for s in scaled]


def _ops():
    from moseq2_detectron_extract_b200.model import ops  # noqa: F401
    return torch.ops.msq


# ---- kernels one by one ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(3, 256, 64, 64), (2, 256, 8, 8), (5, 64, 10, 6)])
def test_group_norm_nhwc_matches_torch(dtype, shape):
    """FPN GroupNorm(32) with the top-down merge riding on it: (GN(x) + up2(top)) * 0.5 against F.group_norm + F.interpolate."""
    msq = _ops()
    g = torch.Generator(device='cuda').manual_seed(shape[2])
    n, c, h, w = shape
    groups = 32 if c % 256 == 0 else 8
    x = (torch.randn(shape, device='cuda', generator=g) * 3 + 1.5).to(dtype).contiguous(memory_format=torch.channels_last)
    top = torch.randn((n, c, h // 2, w // 2), device='cuda', generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    gamma = torch.rand((c,), device='cuda', generator=g) + 0.5
    beta = torch.randn((c,), device='cuda', generator=g)
    want_gn = F.group_norm(x.float(), groups, gamma, beta, 1e-5)
    want = (want_gn + F.interpolate(top.float(), scale_factor=2.0, mode='nearest')) / 2
    got_gn = msq.group_norm_nhwc(x, gamma, beta, groups, 1e-5, None, 1.0)
    got = msq.group_norm_nhwc(x, gamma, beta, groups, 1e-5, top, 0.5)
    assert got.dtype == dtype and got.is_contiguous(memory_format=torch.channels_last)
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7                      # bf16: one rounding of the result
    for a, b in ((got_gn, want_gn), (got, want)):
        assert float((a.float() - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('pooled', [7, 14])
def test_roi_align_v2_matches_detectron2_pooler(dtype, pooled):
    """msq_roi_align_v2 (one launch, in-kernel level assignment, adaptive sampling, aligned=True, channels-last output)
    against detectron2's ROIPooler restated on torchvision.ops.roi_align(aligned=True, sampling_ratio=0)."""
    msq = _ops()
    g = torch.Generator(device='cuda').manual_seed(pooled)
    n_img, ch, size, k = 3, 64, 256, 40
    feats = [torch.randn((n_img, ch, size // s, size // s), device='cuda', generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
             for s in (4, 8, 16, 32)]
    boxes = []
    for i in range(n_img):
        c = torch.rand((k, 2), device='cuda', generator=g) * 240
        wh = torch.exp(torch.rand((k, 2), device='cuda', generator=g) * 5.5)                   # 1 .. 245 px: all four levels
        b = torch.cat([c - wh / 2, c + wh / 2], dim=1).clamp(0, 240)
        b[0] = torch.tensor([0., 0., 240., 240.])                                              # the whole image
        b[1] = torch.tensor([50., 60., 50., 60.])                                              # zero size
        b[2] = torch.tensor([0., 0., 0., 0.])                                                  # a padded proposal slot
        b[3] = torch.tensor([200., 10., 240., 200.])
        b[4] = torch.tensor([10., 20., 130., 150.])                                            # level 3 (112 <= sqrt(area) < 224)
        b[5] = torch.tensor([100., 100., 170., 180.])                                          # level 2
        b[6] = torch.tensor([0., 100., 240., 140.])                                            # elongated: > 12 columns under one bin
        boxes.append(b)
    want = D.roi_pooler([f.float() for f in feats], boxes, pooled)
    got = msq.roi_align_v2(feats, [1 / 4, 1 / 8, 1 / 16, 1 / 32], torch.cat(boxes), k, pooled, 0, 2, 4, 224.0)
    assert got.shape == want.shape and got.dtype == dtype and got.is_contiguous(memory_format=torch.channels_last)
    assert len(set(D.assign_boxes_to_levels(torch.cat(boxes)).tolist())) >= 3
    scale = max(1.0, float(want.abs().max()))
    if dtype == torch.float32:
        # same samples and bilinear weights as torchvision's kernel, summed per row / column instead of per sample (the
        # weights factorise): float32 rounding differs in the last bits
        assert float((got - want).abs().max()) <= 2e-4 * scale
        assert float(((got - want).abs() <= 1e-5 * scale).float().mean()) > 0.99
    else:
        ref = want.to(torch.bfloat16)
        assert float((got == ref).float().mean()) > 0.99
        assert float((got.float() - ref.float()).abs().max()) <= 2 ** -7 * scale


def test_fastrcnn_top1_matches_detectron2_inference():
    """Soft-max + threshold + arg-max + decode + clip against fast_rcnn_inference (score filter, per-class NMS 0.5, top-1)."""
    msq = _ops()
    g = torch.Generator(device='cuda').manual_seed(5)
    n, k = 6, 200
    c = torch.rand((n, k, 2), device='cuda', generator=g) * 240
    wh = torch.rand((n, k, 2), device='cuda', generator=g) * 150 + 2
    props = torch.cat([c - wh / 2, c + wh / 2], dim=-1).clamp(0, 240)
    logits = torch.randn((n * k, 2), device='cuda', generator=g) * 2
    logits[4 * k:5 * k, 0] -= 30                                                  # image 4: nothing above the threshold
    deltas = torch.randn((n * k, 4), device='cuda', generator=g) * 2
    counts = torch.tensor([k, k, 150, k, k, 1], dtype=torch.int32, device='cuda')
    pred = torch.cat([logits, deltas, torch.zeros((n * k, 2), device='cuda')], dim=1)
    box, score, has = msq.fastrcnn_top1(pred, props, counts, 240, 240, 0.05, [10.0, 10.0, 5.0, 5.0])
    cl = counts.tolist()
    want = D.fast_rcnn_inference(torch.cat([logits[i * k:i * k + cl[i]] for i in range(n)]), torch.cat([deltas[i * k:i * k + cl[i]] for i in range(n)]),
                                 [props[i, :cl[i]] for i in range(n)], (240, 240))
    for i, (wb, ws, wc) in enumerate(want):
        assert int(has[i]) == len(wb)
        if len(wb):
            assert torch.allclose(box[i], wb[0], rtol=1e-5, atol=1e-4) and abs(float(score[i]) - float(ws[0])) <= 1e-6
        else:
            assert float(box[i].abs().sum()) == 0 and float(score[i]) == 0
    assert int(has[4]) == 0 and int(has[:4].sum()) == 4


def test_keypoint_decode_matches_detectron2():
    """Batched bicubic arg-max + pooled soft-max score against detectron2's per-RoI heatmaps_to_keypoints loop."""
    msq = _ops()
    g = torch.Generator(device='cuda').manual_seed(1)
    maps = torch.randn((7, 8, 28, 28), device='cuda', generator=g)
    maps = F.avg_pool2d(maps, 5, stride=1, padding=2) * 6                         # smooth: well separated maxima
    rois = torch.tensor([[10.2, 20.7, 90.1, 140.9], [0, 0, 240, 240], [100.5, 50.5, 101.0, 51.0], [5, 5, 35.5, 200],
                         [30, 40, 230.3, 60.8], [0.5, 0.5, 239.5, 239.5], [120, 120, 180, 181]], device='cuda')
    want = D.heatmaps_to_keypoints(maps, rois)[:, :, [0, 1, 3]]
    got = msq.keypoints_from_heatmaps_d2(maps, rois)
    # torch's bicubic kernel fuses multiply-adds we do not: an arg-max may move to a neighbour when two values tie to the last bit
    close = (got[..., :2] - want[..., :2]).abs().amax(dim=-1) <= 1e-3
    assert float(close.float().mean()) >= 0.97
    assert float(((got[..., 2] - want[..., 2]).abs() / want[..., 2])[close].max()) <= 1e-4


@pytest.mark.parametrize('engine', ['fused', 'torch'])
def test_rpn_proposals_match_detectron2(state, engine):
    """find_top_rpn_proposals for a batch (per-level top-k, decode, clip, non-empty, per-level NMS 0.7 via the coordinate trick,
    first 1000 / 100 survivors) against the per-image loop over torchvision.ops.batched_nms, on the same head outputs."""
    msq = _ops()
    from moseq2_detectron_extract_b200.model import ops as _o
    _o.RPN_ENGINE['mode'] = engine          # 'fused': msq_rpn_select (one launch); 'torch': operator by operator
    g = torch.Generator(device='cuda').manual_seed(2)
    feats = [torch.randn((3, 256, s, s), device='cuda', generator=g) * 3 for s in (64, 32, 16, 8, 4)]
    st = dict(state)
    rp = 'proposal_generator.rpn_head.'
    st[rp + 'objectness_logits.weight'] = st[rp + 'objectness_logits.weight'] * 30           # spread the logits
    st[rp + 'anchor_deltas.weight'] = st[rp + 'anchor_deltas.weight'] * 30
    logits, deltas = D.rpn_head(feats, st)
    preds = [torch.cat([lg, dl, torch.zeros_like(lg[:, :1])], dim=1).contiguous(memory_format=torch.channels_last) for lg, dl in zip(logits, deltas)]
    for topk in (1000, 100):
        want = D.rpn_proposals(feats, st, (240, 240), post_nms_topk=topk)
        boxes, scores, counts = msq.rpn_proposals(preds, [4, 8, 16, 32, 64], [32.0, 64.0, 128.0, 256.0, 512.0], [0.5, 1.0, 2.0], 240, 240,
                                                  1000, topk, 0.7)
        assert boxes.shape == (3, topk, 4)
        for i, (wb, ws) in enumerate(want):
            c = int(counts[i])
            assert c == len(wb)
            # the same boxes in the same order (ties in score aside)
            same = (boxes[i, :c] - wb).abs().amax(dim=1) <= 1e-3
            assert float(same.float().mean()) >= 0.99, float(same.float().mean())
            assert torch.allclose(scores[i, :c][same], ws[same], rtol=1e-5, atol=1e-5)
            assert float(boxes[i, c:].abs().sum()) == 0
    # bf16 head outputs (what the graph produces): many equal logits -- both engines must agree on the proposals up to tie order
    preds16 = [p.to(torch.bfloat16) for p in preds]
    _o.RPN_ENGINE['mode'] = 'fused'
    b1, s1, c1 = msq.rpn_proposals(preds16, [4, 8, 16, 32, 64], [32.0, 64.0, 128.0, 256.0, 512.0], [0.5, 1.0, 2.0], 240, 240, 1000, 1000, 0.7)
    _o.RPN_ENGINE['mode'] = 'torch'
    b2, s2, c2 = msq.rpn_proposals(preds16, [4, 8, 16, 32, 64], [32.0, 64.0, 128.0, 256.0, 512.0], [0.5, 1.0, 2.0], 240, 240, 1000, 1000, 0.7)
    _o.RPN_ENGINE['mode'] = 'fused'
    assert (c1 - c2).abs().max() <= 8
    for i in range(3):
        c = int(min(c1[i], c2[i]))
        assert torch.equal(s1[i, :c], s2[i, :c])                        # the same multiset of logits in the same order
        assert float(((b1[i, :c] - b2[i, :c]).abs().amax(dim=1) <= 1e-3).float().mean()) >= 0.9


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('h,w', [(240, 240), (200, 236), (250, 250)])
def test_stem_conv_pool_matches_torch(dtype, h, w):
    """The fused stem (scaling -> normalise -> pad -> 7x7/2 conv + bias -> ReLU -> 3x3/2 max-pool as ONE kernel on the 1-channel
    form) against detector_input -> 3-channel F.conv2d -> relu -> max_pool2d in float32."""
    msq = _ops()
    from moseq2_detectron_extract_b200.model import rcnn
    g = torch.Generator(device='cuda').manual_seed(h)
    chunk = torch.randint(0, 120, (3, h, w), dtype=torch.uint8, device='cuda', generator=g)
    chunk[0, :40] = 0                                                            # a flat region: ReLU / padding paths
    model = rcnn.build_random(seed=1, dtype=torch.float32)
    with torch.no_grad():
        model.stem.bias.normal_(0, 0.5)
        model = rcnn.finalize(model, torch.float32, 'cuda')
        ph, pw = (h + 31) // 32 * 32, (w + 31) // 32 * 32
        x = msq.detector_input(chunk, 0.0, 100.0, True, model.pixel_mean, model.pixel_std, ph, pw, False)
        want = F.max_pool2d(F.relu(F.conv2d(x, model.stem.weight, model.stem.bias, 2, 3)), 3, 2, 1)
        got = msq.stem_conv_pool(chunk, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], ph, pw, model.stem_w49, model.stem_b64,
                                 dtype == torch.bfloat16)
    assert got.shape == want.shape and got.dtype == dtype and got.is_contiguous(memory_format=torch.channels_last)
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7
    assert float((got.float() - want).abs().max()) <= tol * float(want.abs().max())


@pytest.mark.timeout(120)
@pytest.mark.parametrize('h,w', [(240, 240), (200, 236), (250, 250), (64, 48)])
def test_stem_on_tensor_cores_matches_torch(h, w):
    """csrc/stem_tc.cu (im2col written in the swizzle-128B layout + tcgen05.mma + TMEM read-back + pool) against the float32
    reference of the same steps.  Operands are bf16 (normalised input, summed weights), accumulation float32: the result agrees
    with the float32 convolution to bf16 operand rounding (2^-8 relative per product, 49 products)."""
    msq = _ops()
    from moseq2_detectron_extract_b200.model import rcnn
    g = torch.Generator(device='cuda').manual_seed(h + w)
    chunk = torch.randint(0, 120, (5, h, w), dtype=torch.uint8, device='cuda', generator=g)
    chunk[0, : h // 4] = 0
    model = rcnn.build_random(seed=1, dtype=torch.bfloat16)
    with torch.no_grad():
        model.stem.bias.normal_(0, 0.5)
        model = rcnn.finalize(model, torch.bfloat16, 'cuda')
        ph, pw = (h + 31) // 32 * 32, (w + 31) // 32 * 32
        x = msq.detector_input(chunk, 0.0, 100.0, True, model.pixel_mean, model.pixel_std, ph, pw, False)
        w1 = model.stem_w49.t().reshape(64, 1, 7, 7)
        want = F.max_pool2d(F.relu(F.conv2d(x[:, :1], w1, model.stem_b64, 2, 3)), 3, 2, 1)
        got = msq.stem_conv_pool_tc(chunk, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], ph, pw, model.stem_btile, model.stem_b64)
        ref32 = msq.stem_conv_pool(chunk, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], ph, pw, model.stem_w49, model.stem_b64, True)
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == torch.bfloat16 and got.is_contiguous(memory_format=torch.channels_last)
    scale = float(want.abs().max())
    assert float((got.float() - want).abs().max()) <= 0.03 * scale
    assert float((got.float() - want).abs().mean()) <= 0.004 * scale
    assert float((got.float() - ref32.float()).abs().max()) <= 0.03 * scale            # and with the CUDA-core kernel


# ---- the whole graph ------------------------------------------------------------------------------------------------------
def test_graph_float32_matches_detectron2_restatement(state, images):
    """from_detectron2_state_dict (FrozenBN folded, fused epilogues, merged RPN predictor, permuted fc1, channels-last, batched
    heads) in float32 against the eager restatement on the same tensors: pyramid features to 1e-4 of their range, then the same
    detection for every image -- box to 0.05 px, score to 1e-4, mask probabilities to 1e-3, keypoints to 0.5 px (BASELINE)."""
    from moseq2_detectron_extract_b200.model import rcnn
    prep, imgs = images
    model = rcnn.from_detectron2_state_dict(state, dtype=torch.float32)
    with torch.no_grad():
        x = D.preprocess(imgs, state)
        want_feats = D.fpn(D.bottom_up(x, state), state)
        got_feats = model.backbone(x.contiguous(memory_format=torch.channels_last))
        for a, b in zip(got_feats, want_feats):
            assert a.shape == b.shape
            assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max())
        want = D.inference(state, imgs)
        got = model([{'image': im} for im in imgs])
    assert len(got) == len(want) == 4
    for g_, w_ in zip(got, want):
        assert set(g_.keys()) == set(w_.keys())
        assert len(w_['pred_boxes']) == 1 and g_['pred_boxes'].shape == (1, 4)
        assert float((g_['pred_boxes'] - w_['pred_boxes']).abs().max()) <= 0.05
        assert abs(float(g_['scores'][0]) - float(w_['scores'][0])) <= 1e-4
        assert g_['pred_classes'].dtype == torch.int64 and int(g_['pred_classes'][0]) == 0
        assert g_['pred_masks'].shape == (1, 1, 28, 28) and float((g_['pred_masks'] - w_['pred_masks']).abs().max()) <= 1e-3
        assert g_['pred_keypoint_heatmaps'].shape == (1, 8, 28, 28)
        assert float((g_['pred_keypoint_heatmaps'] - w_['pred_keypoint_heatmaps']).abs().max()) <= 1e-3 * float(w_['pred_keypoint_heatmaps'].abs().max())
        assert float((g_['pred_keypoints'][..., :2] - w_['pred_keypoints'][..., :2]).abs().max()) <= 0.5


def test_torchscript_round_trip(tmp_path, state, images):
    """The repo's own TorchScript export (ref model/deploy.py:65-110 contract): save, load through Predictor.from_torchscript,
    same outputs as the eager graph from `forward` (list of dicts) and `forward_dense` (batched), reference-shaped
    Predictor.__call__ and the dense hand-over agree with each other."""
    from moseq2_detectron_extract_b200.model import rcnn
    from moseq2_detectron_extract_b200.model.predict import Predictor
    from moseq2_detectron_extract_b200.proc import scale_raw_frames
    from moseq2_detectron_extract_b200.proc.proc import _gather_instances
    prep, imgs = images
    model = rcnn.from_detectron2_state_dict(state, dtype=torch.bfloat16, post_nms_topk=100)
    path = str(tmp_path / 'model.ts')
    rcnn.export_torchscript(model, path)
    pred = Predictor.from_torchscript(path)
    assert pred.is_torchscript and pred.has_dense_entry and int(pred.model.graph_version) == rcnn.GRAPH_VERSION
    assert pred.model.input_format == 'RGB'
    with torch.no_grad():
        eager = model.forward_dense(prep, 0.0, 100.0, True)
        loaded = pred.model.forward_dense(prep, 0.0, 100.0, True)
        for a, b in zip(eager, loaded):
            assert torch.equal(a, b)
        out_list = pred.model([{'image': im} for im in imgs])
        assert [set(o.keys()) for o in out_list] == [{'pred_boxes', 'scores', 'pred_classes', 'pred_masks', 'pred_keypoints',
                                                      'pred_keypoint_heatmaps'}] * 4
        # forward() normalises in float32 torch ops, forward_dense in the input kernel: same detections
        for i, o in enumerate(out_list):
            assert o['pred_boxes'].shape == (1, 4) and float((o['pred_boxes'][0] - eager[0][i]).abs().max()) <= 0.5
    # reference-shaped call (N, H, W, 1) uint8 numpy vs the dense hand-over
    ref_out = pred(scale_raw_frames(prep.cpu().numpy()[:, :, :, None], 0, 100))
    assert len(ref_out) == 4 and ref_out[0]['instances'].image_size == (240, 240)
    inst = ref_out[0]['instances']
    assert inst.pred_masks.shape == (len(inst), 240, 240) and inst.pred_masks.dtype == torch.bool
    ref_masks, ref_kpts, ref_n = _gather_instances(pred.predict_prepared(prep, 0, 100))
    masks, kpts, ninst = pred.predict_dense(prep, 0, 100)
    assert np.array_equal(ninst.cpu().numpy(), ref_n) and torch.equal(masks, ref_masks)
    assert torch.allclose(kpts, ref_kpts, rtol=0, atol=0, equal_nan=True)


def test_graph_bf16_random_init_runs(images):
    """BASELINE configs[2] set-up: random weights, bf16.  Structure, finiteness, boxes inside the image; with detectron2's 1000
    proposals and with 100."""
    from moseq2_detectron_extract_b200.model.predict import Predictor
    prep, _ = images
    for topk in (1000, 100):
        pred = Predictor.from_random_init(post_nms_topk=topk)
        masks, kpts, ninst = pred.predict_dense(prep, 0, 100)
        assert masks.shape == (4, 240, 240) and masks.dtype == torch.uint8 and kpts.shape == (4, 8, 3)
        assert int(ninst.sum()) >= 1
        ok = ninst > 0
        assert torch.isfinite(kpts[ok]).all() and float(kpts[ok][..., :2].min()) >= 0 and float(kpts[ok][..., :2].max()) <= 240


# ---- kernels kept from the first round, against torchvision ----------------------------------------------------------------
def test_nms_sorted_matches_torchvision():
    torchvision = pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import _dev, _lib
    gen = torch.Generator(device='cuda').manual_seed(0)
    n, K = 3, 400
    xy = torch.rand((n, K, 2), device='cuda', generator=gen) * 200
    wh = torch.rand((n, K, 2), device='cuda', generator=gen) * 60 + 1
    boxes = torch.cat([xy, xy + wh], dim=-1).contiguous()
    valid = (torch.rand((n, K), device='cuda', generator=gen) > 0.1)
    scores = torch.arange(K, 0, -1, device='cuda', dtype=torch.float32)
    keep = torch.empty((n, 50), dtype=torch.int32, device='cuda')
    count = torch.empty((n,), dtype=torch.int32, device='cuda')
    _lib.call('msq_nms_sorted', _dev.ptr(boxes), _dev.ptr(valid.to(torch.uint8).contiguous()), n, K, 0.5, 50, _dev.ptr(keep), _dev.ptr(count),
              _dev.stream())
    for i in range(n):
        idx = torch.nonzero(valid[i])[:, 0]
        ref = idx[torchvision.ops.nms(boxes[i, idx], scores[idx], 0.5)][:50]
        c = int(count[i])
        assert c == len(ref) and torch.equal(keep[i, :c].long(), ref)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_roi_align_levels_matches_torchvision(dtype):
    """msq_roi_align_levels (aligned = false, fixed sampling ratio, (R, C, P, P) output) against torchvision.ops.roi_align."""
    torchvision = pytest.importorskip('torchvision')
    from moseq2_detectron_extract_b200 import _dev, _lib
    g = torch.Generator(device='cuda').manual_seed(4)
    n_img, ch, size, pooled, sampling = 2, 64, 256, 7, 2
    feats = [torch.randn((n_img, ch, size // s, size // s), device='cuda', generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
             for s in (4, 8)]
    c = torch.rand((60, 2), device='cuda', generator=g) * size
    wh = torch.exp(torch.rand((60, 2), device='cuda', generator=g) * 5.0)
    rois = torch.cat([torch.randint(0, n_img, (60, 1), device='cuda', generator=g).float(), c - wh / 2, c + wh / 2], dim=1).contiguous()
    levels = (torch.arange(60, device='cuda') % 2).long()
    out = torch.empty((60, ch, pooled, pooled), dtype=dtype, device='cuda')
    _lib.call('msq_roi_align_levels', (ctypes.c_void_p * 2)(*[f.data_ptr() for f in feats]), (ctypes.c_int * 2)(64, 32), (ctypes.c_int * 2)(64, 32),
              (ctypes.c_float * 2)(0.25, 0.125), 2, ch, int(dtype == torch.bfloat16), _dev.ptr(rois), _dev.ptr(levels), 60, pooled, sampling,
              _dev.ptr(out), _dev.stream())
    want = torch.zeros((60, ch, pooled, pooled), device='cuda')
    for lvl, sc in enumerate((0.25, 0.125)):
        sel = levels == lvl
        want[sel] = torchvision.ops.roi_align(feats[lvl].float(), rois[sel], pooled, sc, sampling, aligned=False)
    scale = max(1.0, float(want.abs().max()))
    if dtype == torch.float32:
        assert float((out - want).abs().max()) <= 2e-4 * scale
    else:
        assert float((out == want.to(torch.bfloat16)).float().mean()) > 0.99


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('h,w,vmax', [(240, 240, 100), (250, 250, 100), (200, 236, 80)])
def test_detector_input_matches_torch(bf16, h, w, vmax):
    """msq_detector_input (a3 scaling + 3 channels + (x - mean) / std + zero padding to a multiple of 32, channels-last) against
    the same steps in torch (ref: proc/proc.py:214-234, model/predict.py:74-77, detectron2 GeneralizedRCNN.preprocess_image)."""
    msq = _ops()
    from moseq2_detectron_extract_b200.proc import scale_raw_frames
    g = torch.Generator(device='cuda').manual_seed(h + w)
    chunk = torch.randint(0, 140, (5, h, w), dtype=torch.uint8, device='cuda', generator=g)
    ph, pw = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    got = msq.detector_input(chunk, 0.0, float(vmax), True, [1.12, 2.0, 3.5], [5.79, 4.0, 7.0], ph, pw, bf16)
    scaled = scale_raw_frames(chunk, 0, vmax).float()
    mean = torch.tensor([1.12, 2.0, 3.5], device='cuda').reshape(1, 3, 1, 1)
    std = torch.tensor([5.79, 4.0, 7.0], device='cuda').reshape(1, 3, 1, 1)
    want = F.pad((scaled[:, None] - mean) / std, (0, pw - w, 0, ph - h))
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    if bf16:
        assert float((got.float() - want.to(torch.bfloat16).float()).abs().max()) <= 2 ** -7 * float(want.abs().max())
    else:
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_upsample2x_bilinear_matches_torch_interpolate():
    """msq_upsample2x_bilinear == F.interpolate(x.float(), scale_factor=2, mode='bilinear', align_corners=False) for the layouts the
    keypoint head produces (bf16 channels-last out of the deconvolution, fp32 dense), edge rows / columns included."""
    import torch.nn.functional as F
    msq = _ops()
    g = torch.Generator().manual_seed(3)
    for shape in ((5, 8, 14, 14), (3, 2, 7, 9), (1, 1, 1, 1), (0, 8, 14, 14)):
        x = torch.randn(shape, generator=g).cuda()
        for t in (x, x.bfloat16(), x.bfloat16().contiguous(memory_format=torch.channels_last), x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)):
            want = F.interpolate(t.float().contiguous(), scale_factor=2.0, mode='bilinear', align_corners=False) if shape[0] else torch.empty((0, shape[1], 2 * shape[2], 2 * shape[3]), device='cuda')
            got = msq.upsample2x_bilinear(t)
            assert got.dtype == torch.float32 and got.shape == want.shape and got.is_contiguous()
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
