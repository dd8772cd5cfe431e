"""A few R-CNN graph calls for profilers: python tools/rcnn_once.py [batch] [topk] [--table]
With --table prints the torch.profiler CUDA-kernel table of one call (what runs besides the dense layers)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200 import synthetic  # noqa: E402
from moseq2_detectron_extract_b200.model.predict import Predictor  # noqa: E402
from moseq2_detectron_extract_b200.proc import prep_raw_frames  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
B = int(args[0]) if args else 250
TOPK = int(args[1]) if len(args) > 1 else 100
geom = synthetic.SessionGeometry()
ch = synthetic.generate_chunk(min(B, 250), seed=3, geom=geom)
prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
prep = prep.repeat((B + len(prep) - 1) // len(prep), 1, 1)[:B].contiguous()
pred = Predictor.from_random_init(post_nms_topk=TOPK)
for _ in range(2):
    pred.predict_dense(prep, 0, 100)
torch.cuda.synchronize()
if '--table' in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        pred.predict_dense(prep, 0, 100)
        torch.cuda.synchronize()
    print(f'one predict_dense call, batch {B}, {TOPK} proposals per image')
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90))
else:
    pred.predict_dense(prep, 0, 100)
    torch.cuda.synchronize()
    print('done')
