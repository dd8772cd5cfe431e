"""Per-layer table of the R-CNN graph's dense layers: cuDNN / cuBLAS vs the tcgen05 implicit GEMM (csrc/conv_tc.cu), CUDA events,
inputs of one 500-frame batch.  python tools/conv_layers.py [batch]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200 import synthetic  # noqa: E402
from moseq2_detectron_extract_b200.model import conv_tc, ops, rcnn  # noqa: E402
from moseq2_detectron_extract_b200.proc import prep_raw_frames  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 500
geom = synthetic.SessionGeometry()
ch = synthetic.generate_chunk(min(B, 250), seed=3, geom=geom)
prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
prep = prep.repeat((B + len(prep) - 1) // len(prep), 1, 1)[:B].contiguous()
model = rcnn.build_random(post_nms_topk=100)
ops.CONV_ENGINE['mode'] = 'auto'
with torch.no_grad():
    model.forward_dense(prep, 0.0, 100.0, True)
shapes = list(ops.engine_choices().keys())
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, iters=5):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                                  # inputs start in HBM, not in L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


print(f'{"layer":58s} {"GFLOP":>8s} {"MB":>8s} {"cudnn ms":>9s} {"tc ms":>8s} {"tc TF/s":>8s} {"tc GB/s":>8s} {"winner":>8s}')
tot = {'cudnn': 0.0, 'tc': 0.0, 'best': 0.0}
g = torch.Generator(device='cuda').manual_seed(0)
for key in shapes:
    if key[0] == 'conv':
        _, xs, ws, stride, pad, has_z, relu, has_b = key
        x = torch.randn(xs, device='cuda', generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(ws, device='cuda', generator=g) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        b = torch.zeros((ws[0],), device='cuda', dtype=torch.bfloat16) if has_b else None
        ho, wo = (xs[2] - 1) // stride + 1, (xs[3] - 1) // stride + 1
        z = torch.randn((xs[0], ws[0], ho, wo), device='cuda', generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if has_z else None
        flops = 2.0 * xs[0] * ho * wo * ws[0] * ws[1] * ws[2] * ws[3]
        nbytes = 2.0 * (x.numel() / (stride * stride) + xs[0] * ho * wo * ws[0] * (2 if has_z else 1) + w.numel())
        t_cd = timeit(lambda: ops._conv2d_cudnn(x, w, b, z, relu, stride, pad))
        t_tc = timeit(lambda: conv_tc.try_conv2d(x, w, b, z, relu, stride, pad)) if conv_tc.try_conv2d(x, w, b, z, relu, stride, pad) is not None else float('nan')
        name = f'conv{ws[2]}x{ws[3]}/{stride} {ws[1]}->{ws[0]} @{xs[2]}x{xs[3]} n={xs[0]}' + (' +res' if has_z else '') + (' relu' if relu else '')
    else:
        _, xs, ws, relu, has_b = key
        x = torch.randn(xs, device='cuda', generator=g).to(torch.bfloat16)
        w = (torch.randn(ws, device='cuda', generator=g) * 0.05).to(torch.bfloat16)
        b = torch.zeros((ws[0],), device='cuda', dtype=torch.bfloat16) if has_b else None
        flops = 2.0 * xs[0] * ws[0] * ws[1]
        nbytes = 2.0 * (x.numel() + xs[0] * ws[0] + w.numel())
        t_cd = timeit(lambda: ops._linear_cublas(x, w, b, relu))
        r = conv_tc.try_linear(x, w, b, relu)
        t_tc = timeit(lambda: conv_tc.try_linear(x, w, b, relu)) if r is not None else float('nan')
        name = f'linear {ws[1]}->{ws[0]} rows={xs[0]}'
    win = 'tcgen05' if t_tc < t_cd else ('cudnn' if t_tc == t_tc else 'cudnn*')     # * = shape not served by conv_tc
    tot['cudnn'] += t_cd; tot['tc'] += t_tc if t_tc == t_tc else t_cd; tot['best'] += min(t_cd, t_tc) if t_tc == t_tc else t_cd
    print(f'{name:58s} {flops / 1e9:8.1f} {nbytes / 1e6:8.1f} {t_cd:9.3f} {t_tc:8.3f} {flops / t_tc / 1e9:8.0f} {nbytes / t_tc / 1e6:8.0f} {win:>8s}')
print(f'sum over distinct layer shapes: cudnn {tot["cudnn"]:.2f} ms, tcgen05 {tot["tc"]:.2f} ms, best-of {tot["best"]:.2f} ms  (repeated blocks counted once)')
