"""Kernel-level table of one R-CNN batch (torch.profiler, CUDA activities): which library kernels the
configs[2] path spends its time in.  Usage: python tools/rcnn_kernels.py [batch] > table.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moseq2_detectron_extract_b200.model.predict import Predictor  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    pred = Predictor.from_random_init(detections_per_img=1, amp=True)
    g = torch.Generator(device='cuda').manual_seed(0)
    chunk = torch.randint(0, 100, (n, 256, 256), dtype=torch.uint8, device='cuda', generator=g)
    for _ in range(3):
        pred.predict_dense(chunk, 0, 100)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=len(sys.argv) > 2) as prof:
        pred.predict_dense(chunk, 0, 100)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90))
    if len(sys.argv) > 2:                                  # second argument: where the many-call ops come from
        rows = [e for e in prof.key_averages(group_by_stack_n=8) if e.count >= n // 2 and e.key.startswith('aten::')]
        for e in sorted(rows, key=lambda e: -e.count)[:25]:
            print(e.count, e.key, ' <- '.join(str(f) for f in e.stack[:8] if 'site-packages/torch/' not in str(f))[:600])


if __name__ == '__main__':
    main()
