"""Session set-up timing: proc.get_roi on the GPU next to the oracle restatement of the reference on the host
(oracle/roi_oracle.py, test infrastructure) on the same synthetic background.  Usage: python tools/roi_times.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import roi_oracle  # noqa: E402
from moseq2_detectron_extract_b200 import _lib, proc  # noqa: E402


def main():
    for h, w in ((424, 512), (576, 640)):
        bg = roi_oracle.synthetic_bground(h=h, w=w, seed=0)
        dev = torch.from_numpy(bg).cuda()
        np.random.seed(0)
        proc.get_roi(dev)
        torch.cuda.synchronize()
        _lib.kernel_timing(True)
        np.random.seed(0)
        t0 = time.perf_counter()
        out = proc.get_roi(dev)
        torch.cuda.synchronize()
        gpu_s = time.perf_counter() - t0
        kern = _lib.kernel_timing_collect()['session_roi']
        _lib.kernel_timing(False)
        np.random.seed(0)
        t0 = time.perf_counter()
        ref = roi_oracle.get_roi(bg)
        cpu_s = time.perf_counter() - t0
        same = all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(out[0], ref[0]))
        print(f'{w}x{h}: {len(out[0])} regions; get_roi {gpu_s * 1e3:.1f} ms wall ({kern[0]:.2f} ms in {kern[1]} timed launches), '
              f'oracle on the host {cpu_s * 1e3:.0f} ms; masks identical: {same}')


if __name__ == '__main__':
    main()
