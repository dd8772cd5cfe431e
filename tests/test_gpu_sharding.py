"""GPU: the sharded session runner (shard.run_session_shard) -- one process per rank driving the real step classes on its
contiguous chunk range -- must reproduce the single-rank result frame for frame, across chunk boundaries and with
`chunk_overlap` (ref: io/util.py:24-35 chunk sequence; pipeline/write_results_step.py:54-73 drops the overlap frames).

Two rank processes are spawned; with fewer than 2 devices both run on cuda:0 (ranks never wait on one another -- there is no
collective on this path -- so sharing a device is safe), with 2 or more each takes its own GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FRAMES, CHUNK, OVERLAP = 190, 40, 6


def _config(geom_cfg, tmp=None):
    cfg = dict(geom_cfg)
    cfg.update(chunk_size=CHUNK, chunk_overlap=OVERLAP, nframes=N_FRAMES, results_to_host=True)
    return cfg


def _worker(rank, world, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch as th
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.shard import run_session_shard
    dev = rank if th.cuda.device_count() >= world else 0
    sess = synthetic.SyntheticSession(N_FRAMES, seed=5, missing_every=17, mask_holes=True)
    res = run_session_shard(sess, _config(synthetic.default_config(sess.geom)), rank=rank, world=world, device_index=dev)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), **res)


@pytest.mark.timeout(600)
def test_two_rank_gpu_shards_equal_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from moseq2_detectron_extract_b200 import synthetic
    from moseq2_detectron_extract_b200.shard import chunk_ranges, concat_results, run_session_shard, shard_chunks
    sess = synthetic.SyntheticSession(N_FRAMES, seed=5, missing_every=17, mask_holes=True)
    single = run_session_shard(sess, _config(synthetic.default_config(sess.geom)))
    assert np.array_equal(single['frame_idxs'], np.arange(N_FRAMES))              # every frame exactly once despite the overlap
    mp.spawn(_worker, args=(2, str(tmp_path)), nprocs=2, join=True)
    parts = [dict(np.load(os.path.join(tmp_path, f'rank{r}.npz'))) for r in range(2)]
    chunks = chunk_ranges(N_FRAMES, CHUNK, OVERLAP)
    for r in range(2):                                                             # contiguous chunk ranges, whole chunks
        mine = shard_chunks(len(chunks), r, 2)
        first = chunks[mine.start].start + (OVERLAP if mine.start > 0 else 0)
        assert parts[r]['frame_idxs'][0] == first and parts[r]['frame_idxs'][-1] == chunks[mine.stop - 1].stop - 1
    merged = concat_results(parts)
    assert np.array_equal(merged['frame_idxs'], np.arange(N_FRAMES))
    assert set(merged.keys()) == set(single.keys()) and 'scalars/velocity_2d_mm' in merged and 'keypoints/reference/Nose_x_px' in merged
    for k in single:                                                               # bit for bit, NaN rows (missing instances) included
        assert np.array_equal(merged[k], single[k], equal_nan=True), k
    # chunk-local semantics survive sharding: velocities restart at each chunk's first frame (incl. its overlap frames)
    assert single['depth_frames'].shape == (N_FRAMES, 80, 80) and single['depth_frames'].any()
