"""Does one predict_dense call capture into a CUDA graph, and what does a replay cost?  python tools/rcnn_graph_probe.py [batch] [topk]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200 import synthetic  # noqa: E402
from moseq2_detectron_extract_b200.model.predict import Predictor  # noqa: E402
from moseq2_detectron_extract_b200.proc import prep_raw_frames  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 500
TOPK = int(sys.argv[2]) if len(sys.argv) > 2 else 100
geom = synthetic.SessionGeometry()
ch = synthetic.generate_chunk(250, seed=3, geom=geom)
prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
prep = prep.repeat((B + len(prep) - 1) // len(prep), 1, 1)[:B].contiguous()
pred = Predictor.from_random_init(post_nms_topk=TOPK, scripted=True)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    t_issue = (time.perf_counter() - t0) / iters * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, t_issue


for _ in range(3):
    ref = pred.predict_dense(prep, 0, 100)
torch.cuda.synchronize()
gpu_ms, issue_ms = timed(lambda: pred.predict_dense(prep, 0, 100))
print(f'eager: {gpu_ms:.2f} ms per call on the device, {issue_ms:.2f} ms of host time to issue it')
static_in = prep.clone()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        pred.predict_dense(static_in, 0, 100)
torch.cuda.current_stream().wait_stream(side)
try:
    with torch.cuda.graph(g):
        out = pred.predict_dense(static_in, 0, 100)
except Exception as e:                                      # pylint: disable=broad-except
    print('capture failed:', type(e).__name__, str(e)[:400])
    sys.exit(0)
g.replay()
torch.cuda.synchronize()
same = all(torch.allclose(a.double(), b.double(), rtol=0, atol=0, equal_nan=True) for a, b in zip(out, ref))
print('replay equals eager:', same)
gpu_ms, issue_ms = timed(g.replay)
print(f'graph replay: {gpu_ms:.2f} ms per call on the device, {issue_ms:.3f} ms of host time')
