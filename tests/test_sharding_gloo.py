"""world_size=2 gloo test (CPU): sharded extraction == single-rank extraction, frame for frame.

The per-chunk work is the oracle here (no GPU in this container); what is under test is the host-side
partition / ordering / gather logic that the GPU runner shares."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FRAMES, CHUNK = 46, 8


def _process_chunk_factory():
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import extract_oracle as O
    from moseq2_detectron_extract_b200 import synthetic
    geom = synthetic.SessionGeometry()
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)

    def process_chunk(ci, frames):
        ch = synthetic.generate_chunk(len(frames), seed=0, geom=geom, t0=frames.start, missing_every=11)
        prep = O.prep_frames(ch.frames, bg, roi, 0, 100)
        res = O.extract_chunk(prep, ch.masks, ch.keypoints, ch.num_instances, 0, 100, 673.0, (80, 80))
        return {'centroid': res['features']['centroid'], 'angle': res['features']['orientation'],
                'velocity_2d_px': res['scalars']['velocity_2d_px'], 'depth_frames': res['depth_frames']}
    return process_chunk


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from moseq2_detectron_extract_b200.shard import ShardedExtractor
    ex = ShardedExtractor(N_FRAMES, CHUNK, _process_chunk_factory(), rank=rank, world=world)
    local = ex.run()
    merged = ex.gather(local)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'merged.npz'), **merged)
    np.save(os.path.join(out_dir, f'owned_{rank}.npy'), local['frame_idxs'])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_extract_equals_single_rank(tmp_path):
    from moseq2_detectron_extract_b200.shard import ShardedExtractor
    single = ShardedExtractor(N_FRAMES, CHUNK, _process_chunk_factory()).run()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(os.path.join(tmp_path, 'merged.npz'))
    owned = [np.load(os.path.join(tmp_path, f'owned_{r}.npy')) for r in range(2)]
    assert owned[0][0] == 0 and owned[0][-1] + 1 == owned[1][0] and owned[1][-1] == N_FRAMES - 1   # contiguous shards
    assert len(owned[0]) % CHUNK == 0                                                              # whole chunks
    assert np.array_equal(merged['frame_idxs'], np.arange(N_FRAMES))
    for k in single:
        assert np.array_equal(merged[k], single[k], equal_nan=True), k
    # velocities restart at every chunk start: chunk-local semantics survive sharding
    assert np.all(merged['velocity_2d_px'][::CHUNK][~np.isnan(merged['velocity_2d_px'][::CHUNK])] == 0)
