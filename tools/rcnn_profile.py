"""Where does the random-init Keypoint+Mask R-CNN path spend its time? (development aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from moseq2_detectron_extract_b200 import synthetic
from moseq2_detectron_extract_b200.model.predict import Predictor
from moseq2_detectron_extract_b200.proc import prep_raw_frames

geom = synthetic.SessionGeometry()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
ch = synthetic.generate_chunk(B, seed=9, geom=geom)
prep = prep_raw_frames(torch.from_numpy(ch.frames).cuda(), bground_im=synthetic.make_background(geom), roi=synthetic.make_roi(geom), vmin=0, vmax=100)
torch.backends.cudnn.benchmark = bool(int(os.environ.get("CUDNN_BENCH", "0")))
pred = Predictor.from_random_init(detections_per_img=1, amp=True)
for _ in range(2):
    pred.predict_dense(prep, 0, 100)
torch.cuda.synchronize()
t = time.time(); pred.predict_dense(prep, 0, 100); torch.cuda.synchronize(); print(f'batch of {B}:', (time.time() - t) * 1e3, 'ms')
model = pred.model.model
import torchvision
from torchvision.models.detection import roi_heads as RH, rpn as RPN
timers = {}
def wrap(obj, name, label):
    fn = getattr(obj, name)
    if isinstance(fn, torch.nn.Module):          # time a sub-module through its forward
        obj, name, fn = fn, 'forward', fn.forward
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.time()
        out = fn(*a, **k)
        torch.cuda.synchronize(); timers[label] = timers.get(label, 0) + (time.time() - t0) * 1e3
        return out
    setattr(obj, name, w)
wrap(model, 'transform', 'transform (normalize/resize/batch)')
wrap(model, 'backbone', 'backbone')
wrap(model.rpn, 'forward', 'rpn total')
wrap(model.rpn, 'filter_proposals', '  rpn.filter_proposals')
wrap(model.roi_heads, 'forward', 'roi_heads total')
wrap(model.roi_heads, 'postprocess_detections', '  roi.postprocess_detections')
wrap(model.roi_heads, 'box_roi_pool', '  roi.box_roi_pool')
wrap(model.roi_heads, 'mask_roi_pool', '  roi.mask_roi_pool')
wrap(model.roi_heads, 'keypoint_roi_pool', '  roi.keypoint_roi_pool')
wrap(model.roi_heads, 'keypoint_head', '  roi.keypoint_head')
wrap(RH, 'keypointrcnn_inference', '  roi.keypointrcnn_inference')
wrap(RH, 'maskrcnn_inference', '  roi.maskrcnn_inference')
orig_post = model.transform.postprocess
torch.cuda.synchronize(); t = time.time(); pred.predict_dense(prep, 0, 100); torch.cuda.synchronize(); print('instrumented batch:', (time.time() - t) * 1e3, 'ms')
for k, v in timers.items():
    print(f'{k:40s} {v:8.1f} ms')
