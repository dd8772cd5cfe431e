#!/bin/bash
# One `ncu --set full` capture per hand-written R-CNN kernel (one launch each, third occurrence), after the same
# command has run clean without the profiler.  Reports land in gpurun_out/.
set -e
mkdir -p gpurun_out
python tools/rcnn_once.py 250 100 > gpurun_out/rcnn_once_plain.log 2>&1
for k in stem_tc_kernel roi_align_v2_kernel rpn_select_kernel gn_apply_kernel nms_mask_kernel keypoint_decode_d2_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$k \
      python tools/rcnn_once.py 250 100 > gpurun_out/ncu_$k.log 2>&1 || echo "ncu $k failed"
done
