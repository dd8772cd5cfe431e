"""Known-answer vectors of detectron2's own unit tests, restated here because detectron2 itself cannot be installed:
tests/layers/test_roi_align.py (ROIAlign on a 5x5 ramp, aligned and not), tests/modeling/test_anchor_generator.py
(DefaultAnchorGenerator, sizes 32 / 64, ratios 0.25 / 1 / 4, stride 4), tests/modeling/test_box2box_transform.py (delta round
trip), structures/keypoints.py heatmaps_to_keypoints (a peak maps back to its pixel centre).  The oracle (oracle/d2_rcnn_oracle.py)
is checked on the CPU; the CUDA kernels behind torch.ops.msq.* against the same numbers under -m gpu."""
import math

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import d2_rcnn_oracle as D  # noqa: E402

RAMP = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
ROI_ALIGNED = torch.tensor([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
ROI_LEGACY = torch.tensor([[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]])
ANCHORS_AT_ORIGIN = torch.tensor([[-32.0, -8.0, 32.0, 8.0], [-16.0, -16.0, 16.0, 16.0], [-8.0, -32.0, 8.0, 32.0],
                                  [-64.0, -16.0, 64.0, 16.0], [-32.0, -32.0, 32.0, 32.0], [-16.0, -64.0, 16.0, 64.0]])


def test_roi_align_ramp_known_answer_oracle():
    import torchvision
    rois = torch.tensor([[0.0, 1.0, 1.0, 3.0, 3.0]])
    assert torch.equal(torchvision.ops.roi_align(RAMP, rois, (4, 4), 1.0, 0, aligned=True)[0, 0], ROI_ALIGNED)
    assert torch.equal(torchvision.ops.roi_align(RAMP, rois, (4, 4), 1.0, 0, aligned=False)[0, 0], ROI_LEGACY)
    # the pooler of the oracle (level assignment + aligned ROIAlign): a 2 px box lands on the stride-4 level
    feats = [torch.nn.functional.interpolate(RAMP, size=(5 * 8 // s, 5 * 8 // s)) for s in (4, 8, 16, 32)]
    got = D.roi_pooler(feats, [torch.tensor([[4.0, 4.0, 12.0, 12.0]])], 4)
    assert got.shape == (1, 1, 4, 4)


def test_anchor_generator_known_answer():
    cells = torch.cat([D.cell_anchors(32, (0.25, 1.0, 4.0)), D.cell_anchors(64, (0.25, 1.0, 4.0))]).float()
    assert torch.allclose(cells, ANCHORS_AT_ORIGIN)
    from moseq2_detectron_extract_b200.model import ops
    mine = torch.cat([ops.cell_anchors(32, (0.25, 1.0, 4.0)), ops.cell_anchors(64, (0.25, 1.0, 4.0))])
    assert torch.allclose(mine, ANCHORS_AT_ORIGIN)
    # detectron2's expected grid for a 1 x 2 feature map, stride 4, offset 0: the cell anchors shifted by (4, 0)
    grid = ops.grid_anchors(1, 2, 4, 32, (0.25, 1.0, 4.0), 'cpu')
    assert torch.allclose(grid[3:], ANCHORS_AT_ORIGIN[:3] + torch.tensor([4.0, 0.0, 4.0, 0.0]))
    assert torch.allclose(grid[:3], ANCHORS_AT_ORIGIN[:3])


def test_box2box_transform_round_trip():
    """test_box2box_transform.py: apply_deltas(get_deltas(src, dst), src) == dst, with the R-CNN head's weights."""
    g = torch.Generator().manual_seed(0)
    weights = (10.0, 10.0, 5.0, 5.0)
    src = torch.rand((10, 4), generator=g) * 100
    src[:, 2:] += src[:, :2] + 1
    dst = torch.rand((10, 4), generator=g) * 100
    dst[:, 2:] += dst[:, :2] + 1

    def get_deltas(a, b):                                      # detectron2 Box2BoxTransform.get_deltas
        aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
        ax, ay = a[:, 0] + 0.5 * aw, a[:, 1] + 0.5 * ah
        bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        bx, by = b[:, 0] + 0.5 * bw, b[:, 1] + 0.5 * bh
        return torch.stack((weights[0] * (bx - ax) / aw, weights[1] * (by - ay) / ah, weights[2] * torch.log(bw / aw),
                            weights[3] * torch.log(bh / ah)), dim=1)
    deltas = get_deltas(src, dst)
    assert torch.allclose(D.apply_deltas(deltas, src, weights), dst, atol=1e-4)
    from moseq2_detectron_extract_b200.model import ops
    assert torch.allclose(ops.apply_deltas(deltas, src, weights), dst, atol=1e-4)
    # the scale clamp: huge dw / dh are cut at log(1000 / 16)
    big = torch.tensor([[0.0, 0.0, 100.0, 100.0]])
    out = ops.apply_deltas(big, torch.tensor([[0.0, 0.0, 10.0, 10.0]]))
    assert math.isclose(float(out[0, 2] - out[0, 0]), 10 * 1000.0 / 16, rel_tol=1e-5)


def test_heatmap_peak_maps_to_its_pixel_centre():
    """heatmaps_to_keypoints on a RoI exactly as large as its heat-map: the arg-max pixel (i, j) comes back as (j + 0.5, i + 0.5)
    plus the RoI offset -- the +0.5 convention detectron2 documents for continuous coordinates."""
    maps = torch.full((1, 2, 28, 28), -5.0)
    maps[0, 0, 7, 19] = 9.0
    maps[0, 1, 20, 3] = 4.0
    rois = torch.tensor([[10.0, 30.0, 38.0, 58.0]])
    out = D.heatmaps_to_keypoints(maps, rois)
    assert torch.allclose(out[0, :, 0], torch.tensor([10 + 19.5, 10 + 3.5])) and torch.allclose(out[0, :, 1], torch.tensor([30 + 7.5, 30 + 20.5]))


@pytest.mark.gpu
def test_roi_align_ramp_known_answer_kernel():
    """msq_roi_align_v2 on detectron2's 5x5 ramp: the aligned=True answers of tests/layers/test_roi_align.py."""
    from moseq2_detectron_extract_b200.model import ops  # noqa: F401
    feat = RAMP.expand(1, 8, 5, 5).contiguous().cuda().contiguous(memory_format=torch.channels_last)
    for dtype, tol in ((torch.float32, 0.0), (torch.bfloat16, 2 ** -4)):
        got = torch.ops.msq.roi_align_v2([feat.to(dtype)], [1.0], torch.tensor([[1.0, 1.0, 3.0, 3.0]], device='cuda'), 1, 4, 0, 2, 2, 224.0)
        assert got.shape == (1, 8, 4, 4)
        for c in range(8):
            assert float((got[0, c].float().cpu() - ROI_ALIGNED).abs().max()) <= tol


@pytest.mark.gpu
def test_keypoint_peak_known_answer_kernel():
    from moseq2_detectron_extract_b200.model import ops  # noqa: F401
    maps = torch.full((1, 2, 28, 28), -5.0)
    maps[0, 0, 7, 19] = 9.0
    maps[0, 1, 20, 3] = 4.0
    out = torch.ops.msq.keypoints_from_heatmaps_d2(maps.cuda(), torch.tensor([[10.0, 30.0, 38.0, 58.0]], device='cuda')).cpu()
    assert torch.allclose(out[0, :, 0], torch.tensor([10 + 19.5, 10 + 3.5])) and torch.allclose(out[0, :, 1], torch.tensor([30 + 7.5, 30 + 20.5]))
