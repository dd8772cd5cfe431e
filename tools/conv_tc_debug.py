"""Development aid: max error of msq_conv_tc against torch for one shape under forced pixel boxes (MSQ_TC_BOX)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from moseq2_detectron_extract_b200.model import conv_tc

def run(n, cin, cout, hw, k, box):
    if box: os.environ['MSQ_TC_BOX'] = box
    else: os.environ.pop('MSQ_TC_BOX', None)
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn((n, cin, hw, hw), device='cuda', generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn((cout, cin, k, k), device='cuda', generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    got = conv_tc.try_conv2d(x, w, None, None, False, 1, k // 2)
    want = F.conv2d(x.float(), w.float(), None, 1, k // 2)
    torch.cuda.synchronize()
    err = (got.float() - want).abs()
    bad = (err > 0.05).float()
    print(f'n={n} cin={cin} cout={cout} hw={hw} k={k} box={box or "auto":9s} max_err={float(err.max()):.4f} bad_frac={float(bad.mean()):.4f} '
          f'bad rows(y)={sorted(set(torch.nonzero(bad.amax(dim=(0,1,3))).flatten().tolist()))[:16]} bad cols(x)={sorted(set(torch.nonzero(bad.amax(dim=(0,1,2))).flatten().tolist()))[:16]} '
          f'bad imgs={sorted(set(torch.nonzero(bad.amax(dim=(1,2,3))).flatten().tolist()))}')

for k in (1, 3):
    for box in ('16,8,1', '16,1,8', '8,16,1', '2,8,8', '16,4,2'):
        run(7, 64, 64, 14, k, box)
run(7, 64, 64, 16, 3, '16,8,1')
run(8, 64, 64, 14, 3, '16,1,8')
run(7, 64, 64, 7, 3, '8,8,2')
run(7, 64, 64, 7, 1, '8,8,2')
print('--- 256 channels')
for box in ('16,8,1', '16,1,8', '8,16,1', '4,4,8'):
    run(7, 256, 256, 14, 3, box)
run(7, 256, 256, 14, 1, '16,8,1')
run(7, 256, 256, 16, 3, '16,8,1')
run(7, 256, 128, 14, 3, '16,8,1')
run(7, 128, 256, 14, 3, '16,8,1')
run(7, 64, 256, 14, 3, '16,8,1')
run(7, 256, 64, 14, 3, '16,8,1')
run(2, 256, 256, 64, 3, None)
run(9, 256, 256, 4, 3, None)
