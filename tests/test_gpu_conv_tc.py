"""GPU: csrc/conv_tc.cu -- the tcgen05 / TMA implicit-GEMM convolution -- against torch (float32 math on the same bf16 operands).
Tolerance: the kernel accumulates in float32 and rounds once to bf16, so results agree with the float32 reference to bf16
rounding of the result (2^-8 relative) plus accumulation-order noise."""
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
F = torch.nn.functional


def _check(got, want):
    scale = float(want.abs().max())
    err = float((got.float() - want).abs().max())
    assert err <= 2 ** -7 * scale + 1e-3, (err, scale)
    assert float(((got.float() - want).abs() <= 2 ** -8 * want.abs() + 2 ** -9 * scale).float().mean()) > 0.99


@pytest.mark.timeout(120)
@pytest.mark.parametrize('n,cin,cout,hw,k,stride,bias,res,relu', [
    (3, 64, 64, 64, 1, 1, True, False, True),          # res2 conv1
    (2, 64, 256, 64, 1, 1, True, True, True),          # res2 conv3 + residual
    (2, 256, 128, 64, 1, 2, True, False, True),        # res3 conv1: stride in the 1x1
    (2, 256, 512, 64, 1, 2, True, False, False),       # res3 shortcut
    (3, 64, 64, 64, 3, 1, True, False, True),          # res2 conv2
    (2, 128, 128, 32, 3, 1, True, False, True),
    (5, 512, 512, 8, 3, 1, True, False, True),         # res5 conv2 (two images per tile)
    (9, 256, 256, 4, 3, 1, True, False, True),         # RPN conv on p6 (eight images per tile, ragged last tile)
    (2, 256, 256, 64, 3, 1, False, False, False),      # FPN output conv: no bias, no activation
    (7, 256, 256, 14, 3, 1, True, False, True),        # mask head: 14x14 RoI maps in 16x8 boxes
    (11, 512, 512, 7, 3, 1, True, False, True),        # keypoint head: 7x7 RoI maps, two per tile
    (3, 2048, 256, 8, 1, 1, False, False, False),      # FPN lateral5: K = 2048
])
def test_conv_tc_matches_torch(n, cin, cout, hw, k, stride, bias, res, relu):
    from moseq2_detectron_extract_b200.model import conv_tc
    g = torch.Generator(device='cuda').manual_seed(n * 1000 + cin + cout + hw)
    x = torch.randn((n, cin, hw, hw), device='cuda', generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((cout, cin, k, k), device='cuda', generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn((cout,), device='cuda', generator=g).to(torch.bfloat16) if bias else None
    ho = (hw - 1) // stride + 1
    z = torch.randn((n, cout, ho, ho), device='cuda', generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if res else None
    got = conv_tc.try_conv2d(x, w, b, z, relu, stride, k // 2)
    assert got is not None and got.shape == (n, cout, ho, ho) and got.dtype == torch.bfloat16
    assert got.is_contiguous(memory_format=torch.channels_last)
    want = F.conv2d(x.float(), w.float(), None if b is None else b.float(), stride, k // 2)
    if z is not None:
        want = want + z.float()
    if relu:
        want = F.relu(want)
    torch.cuda.synchronize()
    _check(got, want)


@pytest.mark.timeout(120)
@pytest.mark.parametrize('rows,kdim,nout,relu', [(1000, 1024, 1024, True), (777, 12544, 1024, True), (130, 64, 64, False)])
def test_linear_tc_matches_torch(rows, kdim, nout, relu):
    from moseq2_detectron_extract_b200.model import conv_tc
    g = torch.Generator(device='cuda').manual_seed(rows)
    x = torch.randn((rows, kdim), device='cuda', generator=g).to(torch.bfloat16)
    w = (torch.randn((nout, kdim), device='cuda', generator=g) / kdim ** 0.5).to(torch.bfloat16)
    b = torch.randn((nout,), device='cuda', generator=g).to(torch.bfloat16)
    got = conv_tc.try_linear(x, w, b, relu)
    want = F.linear(x.float(), w.float(), b.float())
    if relu:
        want = F.relu(want)
    torch.cuda.synchronize()
    assert got is not None and got.shape == (rows, nout)
    _check(got, want)


def test_conv_tc_declines_what_it_does_not_serve():
    from moseq2_detectron_extract_b200.model import conv_tc
    x = torch.zeros((1, 3, 32, 32), device='cuda', dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.zeros((64, 3, 7, 7), device='cuda', dtype=torch.bfloat16)
    assert conv_tc.try_conv2d(x, w, None, None, True, 2, 3) is None                       # the stem: 7x7, 3 channels
    x = torch.zeros((1, 64, 8, 8), device='cuda', dtype=torch.float32).contiguous(memory_format=torch.channels_last)
    assert conv_tc.try_conv2d(x, torch.zeros((64, 64, 1, 1), device='cuda'), None, None, False, 1, 0) is None   # float32 graph


@pytest.mark.timeout(300)
def test_graph_with_tcgen05_convolutions_matches_cudnn_graph():
    """The whole bf16 graph with its convolutions on conv_tc.cu against the same graph on cuDNN: pyramid features to bf16
    accuracy, same detections."""
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.model import ops, rcnn
    from moseq2_detectron_extract_b200.proc import prep_raw_frames
    geom = synthetic.SessionGeometry()
    ch = synthetic.generate_chunk(6, seed=9, geom=geom)
    prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
    model = rcnn.build_random(seed=3, post_nms_topk=100)
    try:
        with torch.no_grad():
            x = torch.ops.msq.stem_conv_pool(prep, 0.0, 100.0, True, model.pixel_mean[0], model.pixel_std[0], 256, 256, model.stem_w49, model.stem_b64, True)
            ops.CONV_ENGINE['mode'] = 'cudnn'
            want = model.pyramid(x)
            ops.CONV_ENGINE['mode'] = 'tcgen05'
            got = model.pyramid(x)
            for a, b in zip(got, want):
                scale = float(b.float().abs().max())
                assert float((a.float() - b.float()).abs().max()) <= 0.05 * scale           # 50 bf16 layers deep
                assert float((a.float() - b.float()).abs().mean()) <= 0.005 * scale
            out_tc = model.forward_dense(prep, 0.0, 100.0, True)
            ops.CONV_ENGINE['mode'] = 'cudnn'
            out_cd = model.forward_dense(prep, 0.0, 100.0, True)
        assert torch.equal(out_tc[2], out_cd[2])
        assert torch.isfinite(out_tc[0]).all() and torch.isfinite(out_tc[4]).all()
    finally:
        ops.CONV_ENGINE['mode'] = 'cudnn'
