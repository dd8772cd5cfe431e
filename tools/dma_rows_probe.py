"""How fast does one strided H2D transfer (cudaMemcpy3DAsync, msq_copy_roi_rows) go as a function of the row width?
python tools/dma_rows_probe.py   -> rows/s and GB/s for several widths / alignments (1000 frames of 512x424 int16, pinned)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200 import _dev, _lib  # noqa: E402

n, H, W = 1000, 424, 512
host = torch.empty((n, H, W), dtype=torch.int16).pin_memory()
host.random_(0, 1000)
h = 240
print('width_px  x0   rows/s(M)   GB/s')
for w, x0 in ((64, 136), (96, 136), (128, 128), (128, 136), (160, 136), (192, 128), (224, 136), (240, 136), (256, 128), (320, 96), (512, 0)):
    dst = torch.empty((n, h, w), dtype=torch.int16, device='cuda')
    for _ in range(2):
        _lib.call('msq_copy_roi_rows', _dev.ptr(host), n, H, W, 92, x0, h, w, _dev.ptr(dst), _dev.stream())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        _lib.call('msq_copy_roi_rows', _dev.ptr(host), n, H, W, 92, x0, h, w, _dev.ptr(dst), _dev.stream())
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f'{w:8d} {x0:4d} {n * h / ms / 1e3:10.1f} {n * h * w * 2 / ms / 1e6:8.1f}')
