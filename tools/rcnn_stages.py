"""Per-stage device times of the R-CNN graph (model/rcnn.py) with CUDA events: python tools/rcnn_stages.py [batch] [topk]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200 import synthetic  # noqa: E402
from moseq2_detectron_extract_b200.model import ops, rcnn  # noqa: E402
from moseq2_detectron_extract_b200.proc import prep_raw_frames  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 250
TOPK = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ENGINE = sys.argv[3] if len(sys.argv) > 3 else 'cudnn'
ops.CONV_ENGINE['mode'] = ENGINE
geom = synthetic.SessionGeometry()
ch = synthetic.generate_chunk(min(B, 250), seed=3, geom=geom)
prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
prep = prep.repeat((B + len(prep) - 1) // len(prep), 1, 1)[:B].contiguous()
model = rcnn.build_random(post_nms_topk=TOPK)
msq = torch.ops.msq
times = {}


def timed(name, fn, iters=3):
    out = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    times[name] = a.elapsed_time(b) / iters
    return out


with torch.no_grad():
    x = timed('detector_input', lambda: msq.detector_input(prep, 0.0, 100.0, True, model.pixel_mean, model.pixel_std, 256, 256, True))
    s = timed('stem', lambda: model.stem(x))
    s = timed('maxpool', lambda: torch.max_pool2d(s, 3, 2, 1))
    c2 = timed('res2', lambda: model.res2(s))
    c3 = timed('res3', lambda: model.res3(c2))
    c4 = timed('res4', lambda: model.res4(c3))
    c5 = timed('res5', lambda: model.res5(c4))
    feats = timed('backbone_total(stem..fpn)', lambda: model.backbone(x))
    times['fpn'] = times['backbone_total(stem..fpn)'] - sum(times[k] for k in ('stem', 'maxpool', 'res2', 'res3', 'res4', 'res5'))
    timed('stem_tc(tcgen05: input+im2col+mma+relu+pool)', lambda: msq.stem_conv_pool_tc(prep, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], 256,
                                                                                      256, model.stem_btile, model.stem_b64))
    timed('fused_stem(input+conv+relu+pool)', lambda: msq.stem_conv_pool(prep, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], 256, 256,
                                                                          model.stem_w49, model.stem_b64, True))
    preds = timed('rpn_head_convs', lambda: [model.rpn_pred(model.rpn_conv(f)) for f in feats])
    props = timed('rpn_proposals(topk+decode+nms)', lambda: msq.rpn_proposals(preds, model.anchor_strides, model.anchor_sizes, model.anchor_ratios,
                                                                             240, 240, model.pre_nms_topk, model.post_nms_topk, 0.7))
    proposals, _, counts = props
    k = proposals.shape[1]
    pooled = timed('box_roi_align', lambda: msq.roi_align_v2(feats[:4], model.pool_scales, proposals.reshape(-1, 4), k, 7, 0, 2, 4, 224.0))
    flat = pooled.permute(0, 2, 3, 1).reshape(B * k, -1)
    f1 = timed('box_fc1', lambda: msq.linear(flat, model.fc1_w, model.fc1_b, True))
    f2 = timed('box_fc2+pred', lambda: msq.linear(msq.linear(f1, model.fc2_w, model.fc2_b, True), model.box_pred_w, model.box_pred_b, False))
    del pooled, flat, f1, f2
    det = timed('box_head_total', lambda: model.box_head(feats, proposals, counts, 240, 240))
    boxes = det[0]
    timed('mask_head', lambda: model.mask_head(feats, boxes))
    timed('keypoint_head', lambda: model.keypoint_head(feats, boxes))
    timed('graph_total(forward_dense)', lambda: model.forward_dense(prep, 0.0, 100.0, True))
    print('mean proposals per image:', float(counts.float().mean()))
tot = times['graph_total(forward_dense)']
if ENGINE == 'auto':
    ch = ops.engine_choices()
    print('auto engine split:', {e: sum(1 for v in ch.values() if v == e) for e in ('tcgen05', 'cudnn')})
    for k, v in ch.items():
        if v == 'tcgen05':
            print('  tcgen05:', k[:5])
print(json.dumps({'batch': B, 'post_nms_topk': TOPK, 'engine': ENGINE, 'ms': {k: round(v, 3) for k, v in times.items()},
                  'frames_per_s_graph_only': B / tot * 1e3}))
