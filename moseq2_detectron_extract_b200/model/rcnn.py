"""The repo's own Keypoint + Mask R-CNN graph, built to the reference's detectron2 configuration.

Reference: model/config.py:21-94 (`get_base_config`) on top of detectron2's COCO-Keypoints/keypoint_rcnn_R_50_FPN_3x.yaml
(Base-Keypoint-RCNN-FPN -> Base-RCNN-FPN), exported by model/deploy.py:65-110 and called by model/predict.py:53-106:

  * GeneralizedRCNN.inference(do_postprocess=False): (x - PIXEL_MEAN) / PIXEL_STD, zero padding to a multiple of 32
  * ResNet-50 (FrozenBN, STRIDE_IN_1X1) -> FPN p2..p6 with FPN.NORM = 'GN' (32 groups), FPN.FUSE_TYPE = 'avg' (:82-83)
  * StandardRPNHead, anchors 32..512 x (0.5, 1, 2), PRE_NMS_TOPK_TEST 1000 per level, NMS 0.7, POST_NMS_TOPK_TEST 1000
  * StandardROIHeads: ROIAlignV2 (sampling_ratio 0) 7x7 -> 2 x FC 1024 -> 1 class + box; SCORE_THRESH_TEST 0.05, NMS 0.5,
    TEST.DETECTIONS_PER_IMAGE = 1 (:75)
  * mask head 14x14 -> 4 conv -> deconv -> 28x28 sigmoid;  keypoint head POOLER_RESOLUTION 7 (:84) -> 8 conv 512 -> deconv ->
    bilinear x2 -> 28x28 heat-maps -> heatmaps_to_keypoints

What differs from detectron2 is only *how* it runs: activations are channels-last in the compute dtype (bf16 on the GPU),
FrozenBN is folded into the convolutions, every conv carries its bias / ReLU / residual in one call, the RPN's two 1x1
predictors are one convolution, the box head's first Linear reads the RoI features in (bin, channel) order (its weight is
permuted once), and all per-image Python loops of detectron2's inference code are batched kernels (model/ops.py).
`from_detectron2_state_dict` performs those rewrites on a detectron2 checkpoint's tensors.

The module is TorchScript-scriptable: `forward` follows the export contract of ref model/deploy.py:91-97 (list of
{'image': CHW} dicts in, list of dicts of tensors out); `forward_dense` is the batched entry the extract pipeline uses.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops as _ops  # noqa: F401  (registers torch.ops.msq.*)

GRAPH_VERSION = 2


class ConvBias(nn.Module):
    """Convolution (+ bias) (+ residual) (+ ReLU) as one call; weight (Cout, Cin, k, k) kept in channels-last memory."""

    def __init__(self, cin: int, cout: int, k: int, stride: int = 1, pad: int = 0, relu: bool = False, bias: bool = True):
        super().__init__()
        self.stride, self.pad, self.relu = stride, pad, relu
        self.weight = nn.Parameter(torch.empty((cout, cin, k, k)).contiguous(memory_format=torch.channels_last), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros((cout,)), requires_grad=False) if bias else None
        nn.init.kaiming_normal_(self.weight, mode='fan_out', nonlinearity='relu')

    def forward(self, x: Tensor, z: Optional[Tensor] = None) -> Tensor:
        return torch.ops.msq.conv2d(x, self.weight, self.bias, z, self.relu, self.stride, self.pad)


class Bottleneck(nn.Module):
    """detectron2 BottleneckBlock with the stride in the first 1x1 convolution (RESNETS.STRIDE_IN_1X1 = True)."""

    def __init__(self, cin: int, cmid: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = ConvBias(cin, cmid, 1, stride=stride, relu=True)
        self.conv2 = ConvBias(cmid, cmid, 3, pad=1, relu=True)
        self.conv3 = ConvBias(cmid, cout, 1, relu=True)                      # ReLU after the residual add
        self.has_shortcut = cin != cout
        self.shortcut = ConvBias(cin, cout, 1, stride=stride) if self.has_shortcut else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        out = self.conv2(self.conv1(x))
        return self.conv3(out, self.shortcut(x))


class ConvGN(nn.Module):
    """FPN convolution without bias followed by GroupNorm(32); the top-down merge rides on the normalisation kernel."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.conv = ConvBias(cin, cout, k, pad=k // 2, bias=False)
        self.gamma = nn.Parameter(torch.ones((cout,)), requires_grad=False)      # float32 whatever the compute dtype
        self.beta = nn.Parameter(torch.zeros((cout,)), requires_grad=False)

    def forward(self, x: Tensor, top: Optional[Tensor], scale: float) -> Tensor:
        return torch.ops.msq.group_norm_nhwc(self.conv(x), self.gamma, self.beta, 32, 1e-5, top, scale)


def _stage(cin: int, cmid: int, cout: int, blocks: int, stride: int) -> nn.Sequential:
    layers = [Bottleneck(cin, cmid, cout, stride)]
    for _ in range(blocks - 1):
        layers.append(Bottleneck(cout, cmid, cout, 1))
    return nn.Sequential(*layers)


class MoseqRCNN(nn.Module):
    def __init__(self, num_keypoints: int = 8, pixel_mean: Tuple[float, float, float] = (1.12, 1.12, 1.12),
                 pixel_std: Tuple[float, float, float] = (5.79, 5.79, 5.79), pre_nms_topk: int = 1000, post_nms_topk: int = 1000,
                 rpn_nms_thresh: float = 0.7, score_thresh: float = 0.05, keypoint_pooler: int = 7, detections_per_image: int = 1):
        super().__init__()
        if detections_per_image != 1:
            raise NotImplementedError('MoseqRCNN serves TEST.DETECTIONS_PER_IMAGE = 1, the reference configuration (model/config.py:75)')
        self.graph_version = GRAPH_VERSION
        self.input_format = 'RGB'                                             # ref: model/config.py:51, model/predict.py:74
        self.num_keypoints = num_keypoints
        self.pixel_mean: List[float] = [float(v) for v in pixel_mean]
        self.pixel_std: List[float] = [float(v) for v in pixel_std]
        self.size_divisibility = 32
        self.pre_nms_topk, self.post_nms_topk, self.rpn_nms_thresh = pre_nms_topk, post_nms_topk, rpn_nms_thresh
        self.score_thresh = score_thresh
        self.keypoint_pooler = keypoint_pooler
        self.anchor_strides: List[int] = [4, 8, 16, 32, 64]
        self.anchor_sizes: List[float] = [32.0, 64.0, 128.0, 256.0, 512.0]
        self.anchor_ratios: List[float] = [0.5, 1.0, 2.0]
        self.box_weights: List[float] = [10.0, 10.0, 5.0, 5.0]
        self.pool_scales: List[float] = [1.0 / 4, 1.0 / 8, 1.0 / 16, 1.0 / 32]
        # ---- bottom-up ResNet-50 ----
        self.stem = ConvBias(3, 64, 7, stride=2, pad=3, relu=True)
        # grey input replicated to 3 channels + channel-uniform mean / std: the stem is a 1-channel convolution (K = 49) and runs
        # fused with the input staging and the max-pool (csrc/prep.cu stem_conv_pool_kernel); buffers filled by finalize()
        self.stem_uniform = len(set(self.pixel_mean)) == 1 and len(set(self.pixel_std)) == 1
        self.register_buffer('stem_w49', torch.zeros((49, 64)))
        self.register_buffer('stem_b64', torch.zeros((64,)))
        self.register_buffer('stem_btile', torch.zeros((4096,), dtype=torch.int16))     # the same weights as a tcgen05 B operand
        self.stem_tc = True                                                    # bf16 graph: stem on the tensor cores (csrc/stem_tc.cu)
        self.res2 = _stage(64, 64, 256, 3, 1)
        self.res3 = _stage(256, 128, 512, 4, 2)
        self.res4 = _stage(512, 256, 1024, 6, 2)
        self.res5 = _stage(1024, 512, 2048, 3, 2)
        # ---- FPN ----
        self.lateral2, self.lateral3 = ConvGN(256, 256, 1), ConvGN(512, 256, 1)
        self.lateral4, self.lateral5 = ConvGN(1024, 256, 1), ConvGN(2048, 256, 1)
        self.output2, self.output3 = ConvGN(256, 256, 3), ConvGN(256, 256, 3)
        self.output4, self.output5 = ConvGN(256, 256, 3), ConvGN(256, 256, 3)
        # ---- RPN head: shared 3x3 conv + ONE 1x1 predictor (3 objectness logits, 3 x 4 deltas, 1 padding channel) ----
        self.rpn_conv = ConvBias(256, 256, 3, pad=1, relu=True)
        self.rpn_pred = ConvBias(256, 16, 1)
        # ---- box head ----
        self.fc1_w = nn.Parameter(torch.empty((1024, 7 * 7 * 256)), requires_grad=False)      # columns in (bin, channel) order
        self.fc1_b = nn.Parameter(torch.zeros((1024,)), requires_grad=False)
        self.fc2_w = nn.Parameter(torch.empty((1024, 1024)), requires_grad=False)
        self.fc2_b = nn.Parameter(torch.zeros((1024,)), requires_grad=False)
        self.box_pred_w = nn.Parameter(torch.zeros((8, 1024)), requires_grad=False)           # rows: fg logit, bg logit, dx, dy, dw, dh, 0, 0
        self.box_pred_b = nn.Parameter(torch.zeros((8,)), requires_grad=False)
        # ---- mask head ----
        self.mask_fcn = nn.Sequential(*[ConvBias(256, 256, 3, pad=1, relu=True) for _ in range(4)])
        self.mask_deconv_w = nn.Parameter(torch.empty((256, 256, 2, 2)), requires_grad=False)
        self.mask_deconv_b = nn.Parameter(torch.zeros((256,)), requires_grad=False)
        self.mask_pred = ConvBias(256, 1, 1)
        # ---- keypoint head ----
        kp = [ConvBias(256, 512, 3, pad=1, relu=True)] + [ConvBias(512, 512, 3, pad=1, relu=True) for _ in range(7)]
        self.kp_fcn = nn.Sequential(*kp)
        self.kp_deconv_w = nn.Parameter(torch.empty((512, num_keypoints, 4, 4)), requires_grad=False)
        self.kp_deconv_b = nn.Parameter(torch.zeros((num_keypoints,)), requires_grad=False)
        self._init_heads()

    def _init_heads(self) -> None:
        nn.init.normal_(self.rpn_conv.weight, std=0.01)
        nn.init.normal_(self.rpn_pred.weight, std=0.01)
        with torch.no_grad():
            self.rpn_pred.weight[15:].zero_()
        nn.init.kaiming_uniform_(self.fc1_w, a=1)
        nn.init.kaiming_uniform_(self.fc2_w, a=1)
        with torch.no_grad():
            self.box_pred_w[:2].normal_(std=0.01)
            self.box_pred_w[2:6].normal_(std=0.001)
        nn.init.kaiming_normal_(self.mask_deconv_w, mode='fan_out', nonlinearity='relu')
        nn.init.normal_(self.mask_pred.weight, std=0.001)
        nn.init.kaiming_normal_(self.kp_deconv_w, mode='fan_out', nonlinearity='relu')

    # ------------------------------------------------------------------------------------------------------------------
    def _compute_bf16(self) -> bool:
        return self.stem.weight.dtype == torch.bfloat16

    def backbone(self, x: Tensor) -> List[Tensor]:
        """x (n, 3, H, W) normalised, padded, channels-last, compute dtype -> [p2, p3, p4, p5, p6]."""
        x = self.stem(x)
        x = torch.max_pool2d(x, [3, 3], [2, 2], [1, 1])
        return self.pyramid(x)

    def pyramid(self, x: Tensor) -> List[Tensor]:
        """x (n, 64, H/4, W/4): the pooled stem output -> [p2, p3, p4, p5, p6]."""
        c2 = self.res2(x)
        c3 = self.res3(c2)
        c4 = self.res4(c3)
        c5 = self.res5(c4)
        i5 = self.lateral5(c5, None, 1.0)
        i4 = self.lateral4(c4, i5, 0.5)
        i3 = self.lateral3(c3, i4, 0.5)
        i2 = self.lateral2(c2, i3, 0.5)
        p5 = self.output5(i5, None, 1.0)
        p4 = self.output4(i4, None, 1.0)
        p3 = self.output3(i3, None, 1.0)
        p2 = self.output2(i2, None, 1.0)
        p6 = torch.max_pool2d(p5, [1, 1], [2, 2], [0, 0])                                    # LastLevelMaxPool
        return [p2, p3, p4, p5, p6]

    def rpn(self, feats: List[Tensor], img_h: int, img_w: int) -> Tuple[Tensor, Tensor, Tensor]:
        preds: List[Tensor] = []
        for f in feats:
            preds.append(self.rpn_pred(self.rpn_conv(f)))
        return torch.ops.msq.rpn_proposals(preds, self.anchor_strides, self.anchor_sizes, self.anchor_ratios, img_h, img_w,
                                           self.pre_nms_topk, self.post_nms_topk, self.rpn_nms_thresh)

    def box_head(self, feats: List[Tensor], proposals: Tensor, counts: Tensor, img_h: int, img_w: int) -> Tuple[Tensor, Tensor, Tensor]:
        n, k = proposals.shape[0], proposals.shape[1]
        x = torch.ops.msq.roi_align_v2(feats[:4], self.pool_scales, proposals.reshape(-1, 4), k, 7, 0, 2, 4, 224.0)
        x = x.permute(0, 2, 3, 1).reshape(n * k, -1)                          # (R, 49 * 256): a view of the channels-last block
        x = torch.ops.msq.linear(x, self.fc1_w, self.fc1_b, True)
        x = torch.ops.msq.linear(x, self.fc2_w, self.fc2_b, True)
        pred = torch.ops.msq.linear(x, self.box_pred_w, self.box_pred_b, False)
        return torch.ops.msq.fastrcnn_top1(pred, proposals, counts, img_h, img_w, self.score_thresh, self.box_weights)

    def mask_head(self, feats: List[Tensor], boxes: Tensor) -> Tensor:
        x = torch.ops.msq.roi_align_v2(feats[:4], self.pool_scales, boxes, 1, 14, 0, 2, 4, 224.0)
        x = self.mask_fcn(x)
        x = torch.relu(torch.conv_transpose2d(x, self.mask_deconv_w, self.mask_deconv_b, [2, 2]))
        return torch.sigmoid(self.mask_pred(x).float())                       # (n, 1, 28, 28)

    def keypoint_head(self, feats: List[Tensor], boxes: Tensor) -> Tuple[Tensor, Tensor]:
        x = torch.ops.msq.roi_align_v2(feats[:4], self.pool_scales, boxes, 1, self.keypoint_pooler, 0, 2, 4, 224.0)
        x = self.kp_fcn(x)
        x = torch.conv_transpose2d(x, self.kp_deconv_w, self.kp_deconv_b, [2, 2], [1, 1])
        # F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) in float32: torch's kernel takes 1-2 ms for these
        # 3 MB (every thread walks all planes); one thread per output pixel, any input strides, takes microseconds
        heat = torch.ops.msq.upsample2x_bilinear(x)
        return torch.ops.msq.keypoints_from_heatmaps_d2(heat, boxes), heat

    def detect(self, x: Tensor, img_h: int, img_w: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
        return self.detect_from_pyramid(self.backbone(x), img_h, img_w)

    def detect_from_pyramid(self, feats: List[Tensor], img_h: int, img_w: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
        proposals, _, counts = self.rpn(feats, img_h, img_w)
        boxes, scores, has = self.box_head(feats, proposals, counts, img_h, img_w)
        soft = self.mask_head(feats, boxes)
        keypoints, heat = self.keypoint_head(feats, boxes)
        return boxes, scores, has, soft, keypoints, heat

    # ---- batched entry of the extract pipeline: prepared uint8 chunk in, first detection of every frame out ---------------
    @torch.jit.export
    def forward_dense(self, chunk_u8: Tensor, vmin: float, vmax: float, int_limits: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
        """chunk_u8 (n, h, w) uint8 as prep_raw_frames leaves it.  Intensity scaling (ref: proc/proc.py:214-234), channel
        replication (ref: model/predict.py:74-77), normalisation and padding happen in one kernel.  Returns
        boxes (n,4), scores (n), has (n) uint8, soft masks (n,1,28,28), keypoints (n,K,3) [x, y, score], heat-maps (n,K,28,28)."""
        h, w = chunk_u8.shape[1], chunk_u8.shape[2]
        d = self.size_divisibility
        ph, pw = (h + d - 1) // d * d, (w + d - 1) // d * d
        if self.stem_uniform and self.stem_tc and self._compute_bf16():
            x = torch.ops.msq.stem_conv_pool_tc(chunk_u8, vmin, vmax, int_limits, self.pixel_mean[0], self.pixel_std[0], ph, pw, self.stem_btile,
                                                self.stem_b64)
            return self.detect_from_pyramid(self.pyramid(x), h, w)
        if self.stem_uniform:
            x = torch.ops.msq.stem_conv_pool(chunk_u8, vmin, vmax, int_limits, self.pixel_mean[0], self.pixel_std[0], ph, pw, self.stem_w49,
                                             self.stem_b64, self._compute_bf16())
            return self.detect_from_pyramid(self.pyramid(x), h, w)
        x = torch.ops.msq.detector_input(chunk_u8, vmin, vmax, int_limits, self.pixel_mean, self.pixel_std, ph, pw, self._compute_bf16())
        return self.detect(x, h, w)

    # ---- export contract of the reference (ref: model/deploy.py:91-97) ------------------------------------------------------
    def forward(self, inputs: List[Dict[str, Tensor]]) -> List[Dict[str, Tensor]]:
        images: List[Tensor] = []
        for i in inputs:
            images.append(i['image'])
        x = torch.stack(images).to(torch.float32)                             # (n, 3, h, w); all images of a call share one size
        h, w = x.shape[2], x.shape[3]
        mean = torch.tensor(self.pixel_mean, dtype=torch.float32, device=x.device).reshape(1, 3, 1, 1)
        std = torch.tensor(self.pixel_std, dtype=torch.float32, device=x.device).reshape(1, 3, 1, 1)
        x = (x - mean) / std
        d = self.size_divisibility
        ph, pw = (h + d - 1) // d * d, (w + d - 1) // d * d
        x = torch.nn.functional.pad(x, [0, pw - w, 0, ph - h])
        x = x.to(self.stem.weight.dtype).contiguous(memory_format=torch.channels_last)
        boxes, scores, has, soft, keypoints, heat = self.detect(x, h, w)
        counts: List[int] = has.to(torch.int64).cpu().tolist()
        out: List[Dict[str, Tensor]] = []
        for i in range(len(counts)):
            c = counts[i]
            out.append({'pred_boxes': boxes[i:i + c], 'scores': scores[i:i + c],
                        'pred_classes': torch.zeros([c], dtype=torch.int64, device=x.device),
                        'pred_masks': soft[i:i + c], 'pred_keypoints': keypoints[i:i + c], 'pred_keypoint_heatmaps': heat[i:i + c]})
        return out


# ---------------------------------------------------------------------------------------------------------------------
# construction helpers (not part of the scripted graph)
# ---------------------------------------------------------------------------------------------------------------------
def build_random(seed: int = 0, dtype: torch.dtype = torch.bfloat16, device: str = 'cuda', **kwargs) -> MoseqRCNN:
    """Random-initialised graph (BASELINE: no network, no checkpoints), detectron2's initialisers approximately."""
    torch.manual_seed(seed)
    model = MoseqRCNN(**kwargs)
    return finalize(model, dtype, device)


def finalize(model: MoseqRCNN, dtype: torch.dtype, device: str) -> MoseqRCNN:
    """Move to `device`, cast every dense-contraction operand to the compute dtype (GroupNorm affine stays float32) and put
    the convolution weights in channels-last memory."""
    model = model.to(device).eval()
    with torch.no_grad():       # 1-channel form of the stem (float32 whatever the compute dtype): weights summed over the input channels
        model.stem_w49 = model.stem.weight.detach().float().sum(dim=1).permute(1, 2, 0).reshape(49, 64).contiguous()
        model.stem_b64 = model.stem.bias.detach().float().contiguous()
        model.stem_btile = stem_b_tile(model.stem_w49).to(model.stem_w49.device)
    for name, p in model.named_parameters():
        if name.endswith('gamma') or name.endswith('beta'):
            p.data = p.data.float().contiguous()
        else:
            t = p.data.to(dtype)
            p.data = t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 and 'deconv' not in name else t.contiguous()
    return model


def stem_b_tile(w49x64: Tensor) -> Tensor:
    """(49, 64) float32 summed stem weight -> the 8 KB B operand of csrc/stem_tc.cu: bf16 [64 n][64 k], K-major rows of 128 bytes
    with the 128-byte swizzle (16-byte chunk index XOR row-in-atom), k >= 49 zero.  Returned as 4096 int16 (bf16 bit patterns)."""
    w = torch.zeros((64, 64), dtype=torch.float32)
    w[:, :49] = w49x64.detach().float().cpu().t()
    bits = w.to(torch.bfloat16).view(torch.int16)                             # [n][k]
    n = torch.arange(64)[:, None].expand(64, 64)
    k = torch.arange(64)[None, :].expand(64, 64)
    pos = n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)                        # in 2-byte units
    tile = torch.zeros((4096,), dtype=torch.int16)
    tile[pos.reshape(-1)] = bits.reshape(-1)
    return tile


def _fold(state: Dict[str, Tensor], name: str) -> Tuple[Tensor, Tensor]:
    """Conv2d + FrozenBatchNorm2d of detectron2 -> (weight, bias): scale = w_bn * rsqrt(var + 1e-5), shift = b_bn - mean * scale."""
    w = state[name + '.weight'].float()
    nw, nb = state[name + '.norm.weight'].float(), state[name + '.norm.bias'].float()
    rm, rv = state[name + '.norm.running_mean'].float(), state[name + '.norm.running_var'].float()
    scale = nw * torch.rsqrt(rv + 1e-5)
    return w * scale.reshape(-1, 1, 1, 1), nb - rm * scale


def from_detectron2_state_dict(state: Dict[str, Tensor], dtype: torch.dtype = torch.bfloat16, device: str = 'cuda', **kwargs) -> MoseqRCNN:
    """Build the graph from the tensors of a detectron2 checkpoint of the reference's configuration (names as
    detectron2's GeneralizedRCNN registers them: backbone.bottom_up.*, backbone.fpn_*, proposal_generator.rpn_head.*,
    roi_heads.{box_head,box_predictor,mask_head,keypoint_head}.*)."""
    state = {k: torch.as_tensor(v) for k, v in state.items()}
    if 'pixel_mean' in state:
        kwargs.setdefault('pixel_mean', tuple(float(v) for v in state['pixel_mean'].reshape(-1)))
        kwargs.setdefault('pixel_std', tuple(float(v) for v in state['pixel_std'].reshape(-1)))
    nk = int(state['roi_heads.keypoint_head.score_lowres.weight'].shape[1])
    model = MoseqRCNN(num_keypoints=nk, **kwargs)

    def put(conv: ConvBias, w: Tensor, b: Optional[Tensor]) -> None:
        conv.weight.data = w.float().contiguous(memory_format=torch.channels_last)
        if b is not None:
            conv.bias.data = b.float().contiguous()

    with torch.no_grad():
        bu = 'backbone.bottom_up.'
        put(model.stem, *_fold(state, bu + 'stem.conv1'))
        for stage_name in ('res2', 'res3', 'res4', 'res5'):
            for i, block in enumerate(getattr(model, stage_name)):
                p = f'{bu}{stage_name}.{i}.'
                put(block.conv1, *_fold(state, p + 'conv1'))
                put(block.conv2, *_fold(state, p + 'conv2'))
                put(block.conv3, *_fold(state, p + 'conv3'))
                if block.has_shortcut:
                    put(block.shortcut, *_fold(state, p + 'shortcut'))
        for lvl in (2, 3, 4, 5):
            for kind in ('lateral', 'output'):
                mod: ConvGN = getattr(model, f'{kind}{lvl}')
                put(mod.conv, state[f'backbone.fpn_{kind}{lvl}.weight'], None)
                mod.gamma.data = state[f'backbone.fpn_{kind}{lvl}.norm.weight'].float().contiguous()
                mod.beta.data = state[f'backbone.fpn_{kind}{lvl}.norm.bias'].float().contiguous()
        rp = 'proposal_generator.rpn_head.'
        put(model.rpn_conv, state[rp + 'conv.weight'], state[rp + 'conv.bias'])
        a = int(state[rp + 'objectness_logits.weight'].shape[0])
        w16 = torch.zeros((16, 256, 1, 1))
        b16 = torch.zeros((16,))
        w16[:a] = state[rp + 'objectness_logits.weight'].float()
        b16[:a] = state[rp + 'objectness_logits.bias'].float()
        w16[a:5 * a] = state[rp + 'anchor_deltas.weight'].float()
        b16[a:5 * a] = state[rp + 'anchor_deltas.bias'].float()
        put(model.rpn_pred, w16, b16)
        bh = 'roi_heads.box_head.'
        # fc1 reads (C, 7, 7)-flattened features in detectron2; here the RoI block is (7, 7, C): permute the columns once
        fc1 = state[bh + 'fc1.weight'].float()
        model.fc1_w.data = fc1.reshape(fc1.shape[0], 256, 7, 7).permute(0, 2, 3, 1).reshape(fc1.shape[0], -1).contiguous()
        model.fc1_b.data = state[bh + 'fc1.bias'].float()
        model.fc2_w.data = state[bh + 'fc2.weight'].float().contiguous()
        model.fc2_b.data = state[bh + 'fc2.bias'].float()
        bp = 'roi_heads.box_predictor.'
        w8, b8 = torch.zeros((8, 1024)), torch.zeros((8,))
        w8[:2], b8[:2] = state[bp + 'cls_score.weight'].float(), state[bp + 'cls_score.bias'].float()        # [class 0, background]
        w8[2:6], b8[2:6] = state[bp + 'bbox_pred.weight'].float(), state[bp + 'bbox_pred.bias'].float()
        model.box_pred_w.data, model.box_pred_b.data = w8, b8
        mh = 'roi_heads.mask_head.'
        for i, conv in enumerate(model.mask_fcn):
            put(conv, state[f'{mh}mask_fcn{i + 1}.weight'], state[f'{mh}mask_fcn{i + 1}.bias'])
        model.mask_deconv_w.data = state[mh + 'deconv.weight'].float().contiguous()
        model.mask_deconv_b.data = state[mh + 'deconv.bias'].float()
        put(model.mask_pred, state[mh + 'predictor.weight'], state[mh + 'predictor.bias'])
        kh = 'roi_heads.keypoint_head.'
        for i, conv in enumerate(model.kp_fcn):
            put(conv, state[f'{kh}conv_fcn{i + 1}.weight'], state[f'{kh}conv_fcn{i + 1}.bias'])
        model.kp_deconv_w.data = state[kh + 'score_lowres.weight'].float().contiguous()
        model.kp_deconv_b.data = state[kh + 'score_lowres.bias'].float()
    return finalize(model, dtype, device)


def export_torchscript(model: MoseqRCNN, path: str) -> torch.jit.ScriptModule:
    """The repo's own `model.ts` (what ref model/deploy.py:65-110 produces with detectron2's scripting_with_instances)."""
    scripted = torch.jit.script(model)
    torch.jit.save(scripted, path)
    return scripted


def dense_flops_per_frame(h: int, w: int, proposals: int, num_keypoints: int = 8, keypoint_pooler: int = 7) -> float:
    """FLOPs (2 x multiply-accumulates) of every convolution / Linear the graph executes for one (h, w) frame: the 1-channel stem
    (K = 49), res2..res5 with the stride in the first 1x1, FPN laterals / outputs, the RPN head on p2..p6, the box head on
    `proposals` RoIs, mask and keypoint heads on one detection.  The tensor-pipe roofline of bench.py divides by this."""
    ph, pw = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    macs = 0.0
    ch, cw = (ph - 1) // 2 + 1, (pw - 1) // 2 + 1
    macs += ch * cw * 64 * 49
    fh, fw = (ch - 1) // 2 + 1, (cw - 1) // 2 + 1
    cin = 64
    sizes = []
    for blocks, mid, cout, stride in ((3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)):
        for i in range(blocks):
            if i == 0:
                fh, fw = (fh - 1) // stride + 1, (fw - 1) // stride + 1
                macs += fh * fw * cin * cout                                   # shortcut
            macs += fh * fw * (cin * mid + 9 * mid * mid + mid * cout)
            cin = cout
        sizes.append((fh, fw, cout))
    for fh_, fw_, c in sizes:
        macs += fh_ * fw_ * (c * 256 + 9 * 256 * 256)                          # lateral + output
    levels = [(a, b) for a, b, _ in sizes] + [((sizes[-1][0] - 1) // 2 + 1, (sizes[-1][1] - 1) // 2 + 1)]
    for a, b in levels:
        macs += a * b * (9 * 256 * 256 + 256 * 16)                             # RPN head
    macs += proposals * (49 * 256 * 1024 + 1024 * 1024 + 1024 * 8)             # box head
    macs += 4 * 196 * 9 * 256 * 256 + 196 * 4 * 256 * 256 + 784 * 256          # mask head
    p2 = keypoint_pooler * keypoint_pooler
    macs += p2 * 9 * 256 * 512 + 7 * p2 * 9 * 512 * 512 + p2 * 16 * 512 * num_keypoints
    return 2.0 * macs
