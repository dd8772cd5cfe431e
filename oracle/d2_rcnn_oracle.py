"""TEST INFRASTRUCTURE ONLY -- plain PyTorch (float32, eager, NCHW) restatement of the detectron2 graph the reference
configures and exports.  Only tests/ may import this file; the product (moseq2_detectron_extract_b200/model/rcnn.py) must
never route through it.

**Parity unpinned**: detectron2 is not installed here (the reference's README installs git HEAD; setup.py does not list
it), so nothing below ever ran against the library.  It follows detectron2's published inference code path function by
function, with the arithmetic done by the same PyTorch / torchvision operators detectron2 itself calls:

  GeneralizedRCNN.inference(do_postprocess=False)      detectron2/modeling/meta_arch/rcnn.py
  ResNet BasicStem / BottleneckBlock (STRIDE_IN_1X1)    detectron2/modeling/backbone/resnet.py
  FPN (norm='GN', fuse_type='avg'), LastLevelMaxPool    detectron2/modeling/backbone/fpn.py
  StandardRPNHead, DefaultAnchorGenerator               detectron2/modeling/proposal_generator/rpn.py, anchor_generator.py
  find_top_rpn_proposals                                detectron2/modeling/proposal_generator/proposal_utils.py
  Box2BoxTransform.apply_deltas                         detectron2/modeling/box_regression.py
  ROIPooler + assign_boxes_to_levels, ROIAlign(aligned) detectron2/modeling/poolers.py -> torchvision.ops.roi_align
  FastRCNNConvFCHead, FastRCNNOutputLayers.inference    detectron2/modeling/roi_heads/{box_head,fast_rcnn}.py
  MaskRCNNConvUpsampleHead, mask_rcnn_inference         detectron2/modeling/roi_heads/mask_head.py
  KRCNNConvDeconvUpsampleHead, keypoint_rcnn_inference  detectron2/modeling/roi_heads/keypoint_head.py
  heatmaps_to_keypoints                                 detectron2/structures/keypoints.py

with the configuration of ref model/config.py:21-94 (1 class, GN FPN with avg fusion, keypoint pooler 7, one detection per
image, COCO-Keypoints/keypoint_rcnn_R_50_FPN_3x.yaml defaults otherwise).  The state dict uses detectron2's parameter names,
so `model.rcnn.from_detectron2_state_dict` is exercised on the same tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F
import torchvision

STAGES = (('res2', 3, 64, 256, 1), ('res3', 4, 128, 512, 2), ('res4', 6, 256, 1024, 2), ('res5', 3, 512, 2048, 2))
ANCHOR_SIZES = (32.0, 64.0, 128.0, 256.0, 512.0)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)
STRIDES = (4, 8, 16, 32, 64)
SCALE_CLAMP = math.log(1000.0 / 16)


# ---------------------------------------------------------------------------------------------------------------------
def make_random_state(seed: int = 0, num_keypoints: int = 8, keypoint_pooler: int = 7) -> Dict[str, torch.Tensor]:
    """Random tensors under detectron2's names; FrozenBN statistics and affine terms are random too so that folding matters."""
    g = torch.Generator().manual_seed(seed)
    st: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, bias=False, std=None):
        fan_out = cout * k * k
        std = math.sqrt(2.0 / fan_out) if std is None else std
        st[name + '.weight'] = torch.randn((cout, cin, k, k), generator=g) * std
        if bias:
            st[name + '.bias'] = torch.randn((cout,), generator=g) * 0.01

    def bn(name, c):
        st[name + '.norm.weight'] = 1.0 + 0.1 * torch.randn((c,), generator=g)
        st[name + '.norm.bias'] = 0.05 * torch.randn((c,), generator=g)
        st[name + '.norm.running_mean'] = 0.05 * torch.randn((c,), generator=g)
        st[name + '.norm.running_var'] = 1.0 + 0.2 * torch.rand((c,), generator=g)

    bu = 'backbone.bottom_up.'
    conv(bu + 'stem.conv1', 64, 3, 7); bn(bu + 'stem.conv1', 64)
    cin = 64
    for name, blocks, mid, cout, _ in STAGES:
        for i in range(blocks):
            p = f'{bu}{name}.{i}.'
            if i == 0:
                conv(p + 'shortcut', cout, cin, 1); bn(p + 'shortcut', cout)
            conv(p + 'conv1', mid, cin, 1); bn(p + 'conv1', mid)
            conv(p + 'conv2', mid, mid, 3); bn(p + 'conv2', mid)
            conv(p + 'conv3', cout, mid, 1); bn(p + 'conv3', cout)
            st[p + 'conv3.norm.weight'] *= 0.3                       # keep the residual stream tame over 16 blocks
            cin = cout
    for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        for kind, k, ci in (('lateral', 1, c), ('output', 3, 256)):
            n = f'backbone.fpn_{kind}{lvl}'
            conv(n, 256, ci, k)
            st[n + '.norm.weight'] = 1.0 + 0.1 * torch.randn((256,), generator=g)
            st[n + '.norm.bias'] = 0.05 * torch.randn((256,), generator=g)
    rp = 'proposal_generator.rpn_head.'
    conv(rp + 'conv', 256, 256, 3, bias=True, std=0.01)
    conv(rp + 'objectness_logits', 3, 256, 1, bias=True, std=0.01)
    conv(rp + 'anchor_deltas', 12, 256, 1, bias=True, std=0.01)
    bh = 'roi_heads.box_head.'
    st[bh + 'fc1.weight'] = torch.randn((1024, 256 * 7 * 7), generator=g) * math.sqrt(1.0 / (256 * 49))
    st[bh + 'fc1.bias'] = torch.randn((1024,), generator=g) * 0.01
    st[bh + 'fc2.weight'] = torch.randn((1024, 1024), generator=g) * math.sqrt(1.0 / 1024)
    st[bh + 'fc2.bias'] = torch.randn((1024,), generator=g) * 0.01
    bp = 'roi_heads.box_predictor.'
    st[bp + 'cls_score.weight'] = torch.randn((2, 1024), generator=g) * 0.05
    st[bp + 'cls_score.bias'] = torch.randn((2,), generator=g) * 0.01
    st[bp + 'bbox_pred.weight'] = torch.randn((4, 1024), generator=g) * 0.01
    st[bp + 'bbox_pred.bias'] = torch.randn((4,), generator=g) * 0.01
    mh = 'roi_heads.mask_head.'
    for i in range(4):
        conv(f'{mh}mask_fcn{i + 1}', 256, 256, 3, bias=True)
    st[mh + 'deconv.weight'] = torch.randn((256, 256, 2, 2), generator=g) * math.sqrt(2.0 / (256 * 4))
    st[mh + 'deconv.bias'] = torch.randn((256,), generator=g) * 0.01
    conv(mh + 'predictor', 1, 256, 1, bias=True, std=0.05)
    kh = 'roi_heads.keypoint_head.'
    for i in range(8):
        conv(f'{kh}conv_fcn{i + 1}', 512, 256 if i == 0 else 512, 3, bias=True)
    st[kh + 'score_lowres.weight'] = torch.randn((512, num_keypoints, 4, 4), generator=g) * math.sqrt(2.0 / (512 * 16))
    st[kh + 'score_lowres.bias'] = torch.randn((num_keypoints,), generator=g) * 0.01
    st['pixel_mean'] = torch.tensor([1.12, 1.12, 1.12]).reshape(3, 1, 1)
    st['pixel_std'] = torch.tensor([5.79, 5.79, 5.79]).reshape(3, 1, 1)
    return st


# ---------------------------------------------------------------------------------------------------------------------
def conv_frozen_bn(x, st, name, stride=1, padding=0):
    x = F.conv2d(x, st[name + '.weight'], None, stride, padding)
    scale = st[name + '.norm.weight'] * (st[name + '.norm.running_var'] + 1e-5).rsqrt()
    bias = st[name + '.norm.bias'] - st[name + '.norm.running_mean'] * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def bottom_up(x, st) -> Dict[str, torch.Tensor]:
    bu = 'backbone.bottom_up.'
    x = F.relu_(conv_frozen_bn(x, st, bu + 'stem.conv1', 2, 3))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    out = {}
    for name, blocks, _, _, first_stride in STAGES:
        for i in range(blocks):
            p = f'{bu}{name}.{i}.'
            stride = first_stride if i == 0 else 1
            y = F.relu_(conv_frozen_bn(x, st, p + 'conv1', stride, 0))            # STRIDE_IN_1X1
            y = F.relu_(conv_frozen_bn(y, st, p + 'conv2', 1, 1))
            y = conv_frozen_bn(y, st, p + 'conv3')
            shortcut = conv_frozen_bn(x, st, p + 'shortcut', stride, 0) if (p + 'shortcut.weight') in st else x
            x = F.relu_(y + shortcut)
        out[name] = x
    return out


def fpn(feats, st) -> List[torch.Tensor]:
    def conv_gn(x, name, k):
        y = F.conv2d(x, st[name + '.weight'], None, 1, k // 2)
        return F.group_norm(y, 32, st[name + '.norm.weight'], st[name + '.norm.bias'], 1e-5)

    prev = conv_gn(feats['res5'], 'backbone.fpn_lateral5', 1)
    results = [conv_gn(prev, 'backbone.fpn_output5', 3)]
    for lvl in (4, 3, 2):
        top_down = F.interpolate(prev, scale_factor=2.0, mode='nearest')
        lateral = conv_gn(feats[f'res{lvl}'], f'backbone.fpn_lateral{lvl}', 1)
        prev = (lateral + top_down) / 2                                           # fuse_type == 'avg'
        results.insert(0, conv_gn(prev, f'backbone.fpn_output{lvl}', 3))
    results.append(F.max_pool2d(results[-1], kernel_size=1, stride=2, padding=0))  # LastLevelMaxPool: p6 from p5
    return results                                                                # p2 .. p6


def cell_anchors(size, ratios):
    out = []
    for ar in ratios:
        area = size ** 2.0
        w = math.sqrt(area / ar)
        h = ar * w
        out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out)


def grid_anchors(gh, gw, stride, size, device):
    sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32, device=device)
    sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(sy, sx, indexing='ij')
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors(size, ANCHOR_RATIOS).to(device).view(1, -1, 4)).reshape(-1, 4)


def apply_deltas(deltas, boxes, weights):
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    widths, heights = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    ctr_x, ctr_y = boxes[:, 0] + 0.5 * widths, boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=SCALE_CLAMP)
    pcx, pcy = dx * widths[:, None] + ctr_x[:, None], dy * heights[:, None] + ctr_y[:, None]
    pw, ph = torch.exp(dw) * widths[:, None], torch.exp(dh) * heights[:, None]
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1).reshape(deltas.shape)


def rpn_head(features, st):
    rp = 'proposal_generator.rpn_head.'
    logits, deltas = [], []
    for x in features:
        t = F.relu(F.conv2d(x, st[rp + 'conv.weight'], st[rp + 'conv.bias'], 1, 1))
        logits.append(F.conv2d(t, st[rp + 'objectness_logits.weight'], st[rp + 'objectness_logits.bias']))
        deltas.append(F.conv2d(t, st[rp + 'anchor_deltas.weight'], st[rp + 'anchor_deltas.bias']))
    return logits, deltas


def rpn_proposals(features, st, image_size, pre_nms_topk=1000, post_nms_topk=1000, nms_thresh=0.7):
    """-> per image (boxes (k,4), logits (k))."""
    logits, deltas = rpn_head(features, st)
    n = features[0].shape[0]
    device = features[0].device
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(n, device=device)
    for lvl, (lg, dl) in enumerate(zip(logits, deltas)):
        gh, gw = lg.shape[-2:]
        anchors = grid_anchors(gh, gw, STRIDES[lvl], ANCHOR_SIZES[lvl], device)
        lg = lg.permute(0, 2, 3, 1).flatten(1)
        dl = dl.view(n, -1, 4, gh, gw).permute(0, 3, 4, 1, 2).flatten(1, -2)
        proposals = apply_deltas(dl.reshape(-1, 4), anchors[None].expand(n, -1, -1).reshape(-1, 4), (1.0, 1.0, 1.0, 1.0)).view(n, -1, 4)
        k = min(lg.shape[1], pre_nms_topk)
        sc, idx = lg.topk(k, dim=1)
        topk_scores.append(sc)
        topk_proposals.append(proposals[batch_idx[:, None], idx])
        level_ids.append(torch.full((k,), lvl, dtype=torch.int64, device=device))
    topk_scores, topk_proposals, level_ids = torch.cat(topk_scores, 1), torch.cat(topk_proposals, 1), torch.cat(level_ids)
    h, w = image_size
    results = []
    for i in range(n):
        boxes, scores, lvl = topk_proposals[i], topk_scores[i], level_ids
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
        boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
        boxes = torch.stack((boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h), boxes[:, 2].clamp(0, w), boxes[:, 3].clamp(0, h)), dim=1)
        keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = torchvision.ops.batched_nms(boxes, scores, lvl, nms_thresh)[:post_nms_topk]
        results.append((boxes[keep], scores[keep]))
    return results


def assign_boxes_to_levels(boxes, min_level=2, max_level=5, canonical_box_size=224, canonical_level=4):
    sizes = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lv = torch.floor(canonical_level + torch.log2(sizes / canonical_box_size + 1e-8))
    return torch.clamp(lv, min=min_level, max=max_level).to(torch.int64) - min_level


def roi_pooler(features, box_lists, output_size, sampling_ratio=0):
    """detectron2 ROIPooler (ROIAlignV2) over p2..p5; box_lists: per-image (k_i,4)."""
    rois = torch.cat([torch.cat((torch.full((len(b), 1), i, dtype=b.dtype, device=b.device), b), dim=1) for i, b in enumerate(box_lists)])
    levels = assign_boxes_to_levels(rois[:, 1:])
    out = torch.zeros((len(rois), features[0].shape[1], output_size, output_size), dtype=features[0].dtype, device=features[0].device)
    for lvl in range(4):
        inds = torch.nonzero(levels == lvl).squeeze(1)
        if len(inds):
            out[inds] = torchvision.ops.roi_align(features[lvl], rois[inds], output_size, 1.0 / STRIDES[lvl], sampling_ratio, aligned=True)
    return out


def box_head_outputs(features, st, proposal_boxes):
    x = roi_pooler(features, proposal_boxes, 7)
    bh, bp = 'roi_heads.box_head.', 'roi_heads.box_predictor.'
    x = torch.flatten(x, start_dim=1)
    x = F.relu(F.linear(x, st[bh + 'fc1.weight'], st[bh + 'fc1.bias']))
    x = F.relu(F.linear(x, st[bh + 'fc2.weight'], st[bh + 'fc2.bias']))
    return F.linear(x, st[bp + 'cls_score.weight'], st[bp + 'cls_score.bias']), F.linear(x, st[bp + 'bbox_pred.weight'], st[bp + 'bbox_pred.bias'])


def fast_rcnn_inference(scores_logits, deltas, proposal_boxes, image_size, score_thresh=0.05, nms_thresh=0.5, topk=1):
    counts = [len(b) for b in proposal_boxes]
    boxes = apply_deltas(deltas, torch.cat(proposal_boxes), (10.0, 10.0, 5.0, 5.0))
    probs = F.softmax(scores_logits, dim=-1)
    h, w = image_size
    out = []
    for b, s in zip(boxes.split(counts), probs.split(counts)):
        valid = torch.isfinite(b).all(dim=1) & torch.isfinite(s).all(dim=1)
        b, s = b[valid], s[valid][:, :-1]                                         # the background column is the last one
        b = torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), dim=1).view(-1, 1, 4)
        mask = s > score_thresh
        inds = mask.nonzero()
        bb, ss = b[mask], s[mask]
        keep = torchvision.ops.batched_nms(bb, ss, inds[:, 1], nms_thresh)[:topk]
        out.append((bb[keep], ss[keep], inds[keep, 1]))
    return out


def mask_head(features, st, det_boxes):
    mh = 'roi_heads.mask_head.'
    x = roi_pooler(features, det_boxes, 14)
    for i in range(4):
        x = F.relu(F.conv2d(x, st[f'{mh}mask_fcn{i + 1}.weight'], st[f'{mh}mask_fcn{i + 1}.bias'], 1, 1))
    x = F.relu(F.conv_transpose2d(x, st[mh + 'deconv.weight'], st[mh + 'deconv.bias'], stride=2))
    return F.conv2d(x, st[mh + 'predictor.weight'], st[mh + 'predictor.bias']).sigmoid()


def heatmaps_to_keypoints(maps, rois):
    offset_x, offset_y = rois[:, 0], rois[:, 1]
    widths, heights = (rois[:, 2] - rois[:, 0]).clamp(min=1), (rois[:, 3] - rois[:, 1]).clamp(min=1)
    widths_ceil, heights_ceil = widths.ceil(), heights.ceil()
    num_rois, num_keypoints = maps.shape[:2]
    xy_preds = maps.new_zeros(rois.shape[0], num_keypoints, 4)
    width_corrections, height_corrections = widths / widths_ceil, heights / heights_ceil
    keypoints_idx = torch.arange(num_keypoints, device=maps.device)
    for i in range(num_rois):
        outsize = (int(heights_ceil[i]), int(widths_ceil[i]))
        roi_map = F.interpolate(maps[[i]], size=outsize, mode='bicubic', align_corners=False)
        roi_map = roi_map.reshape(roi_map.shape[1:])
        max_score, _ = roi_map.view(num_keypoints, -1).max(1)
        max_score = max_score.view(num_keypoints, 1, 1)
        tmp_full_resolution = (roi_map - max_score).exp_()
        tmp_pool_resolution = (maps[i] - max_score).exp_()
        roi_map_scores = tmp_full_resolution / tmp_pool_resolution.sum((1, 2), keepdim=True)
        w = roi_map.shape[2]
        pos = roi_map.view(num_keypoints, -1).argmax(1)
        x_int = pos % w
        y_int = (pos - x_int) // w
        x = (x_int.float() + 0.5) * width_corrections[i]
        y = (y_int.float() + 0.5) * height_corrections[i]
        xy_preds[i, :, 0] = x + offset_x[i]
        xy_preds[i, :, 1] = y + offset_y[i]
        xy_preds[i, :, 2] = roi_map[keypoints_idx, y_int, x_int]
        xy_preds[i, :, 3] = roi_map_scores[keypoints_idx, y_int, x_int]
    return xy_preds


def keypoint_head(features, st, det_boxes, pooler_resolution=7):
    kh = 'roi_heads.keypoint_head.'
    x = roi_pooler(features, det_boxes, pooler_resolution)
    for i in range(8):
        x = F.relu(F.conv2d(x, st[f'{kh}conv_fcn{i + 1}.weight'], st[f'{kh}conv_fcn{i + 1}.bias'], 1, 1))
    x = F.conv_transpose2d(x, st[kh + 'score_lowres.weight'], st[kh + 'score_lowres.bias'], stride=2, padding=1)
    heat = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    res = heatmaps_to_keypoints(heat, torch.cat(det_boxes))
    return res[:, :, [0, 1, 3]], heat


def preprocess(images: List[torch.Tensor], st) -> torch.Tensor:
    x = torch.stack([(im.float() - st['pixel_mean']) / st['pixel_std'] for im in images])
    h, w = x.shape[-2:]
    ph, pw = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    return F.pad(x, (0, pw - w, 0, ph - h))


@torch.no_grad()
def inference(st: Dict[str, torch.Tensor], images: List[torch.Tensor], post_nms_topk: int = 1000, keypoint_pooler: int = 7):
    """GeneralizedRCNN.inference(do_postprocess=False) -> per image dict of the fields ref model/deploy.py:73-82 names."""
    h, w = images[0].shape[-2:]
    x = preprocess(images, st)
    feats = fpn(bottom_up(x, st), st)
    props = rpn_proposals(feats, st, (h, w), post_nms_topk=post_nms_topk)
    logits, deltas = box_head_outputs(feats[:4], st, [p[0] for p in props])
    dets = fast_rcnn_inference(logits, deltas, [p[0] for p in props], (h, w))
    det_boxes = [d[0] for d in dets]
    masks = mask_head(feats[:4], st, det_boxes)
    kpts, heat = keypoint_head(feats[:4], st, det_boxes, keypoint_pooler)
    counts = [len(b) for b in det_boxes]
    out = []
    for d, m, k, hm in zip(dets, masks.split(counts), kpts.split(counts), heat.split(counts)):
        out.append({'pred_boxes': d[0], 'scores': d[1], 'pred_classes': d[2], 'pred_masks': m, 'pred_keypoints': k,
                    'pred_keypoint_heatmaps': hm})
    return out
