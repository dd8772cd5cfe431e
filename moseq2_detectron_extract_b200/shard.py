"""Chunk sharding across GPUs.

The independent unit of the extract path is a chunk of `chunk_size` consecutive frames: with
use_tracking=False and <=1 instance per frame there is no state across chunks (velocities restart per chunk,
ref: proc/scalars.py:105-107; the angle filter is chunk-local, ref: proc/proc.py:837).  GPU g of G owns the
contiguous chunk range [g*ceil(n/G), (g+1)*ceil(n/G)); background and ROI are replicated.  There is NO
collective on the data path: every rank extracts its chunks and keeps (or writes) its results; an optional
control-plane gather of the per-frame tables to rank 0 uses torch.distributed object gather.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/.../local_cpulist)."""
    cpus: List[int] = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[Dict[str, object]]:
    """Pin this process to the CPUs local to GPU `device_index` (one process per GPU): page-locked buffers allocated
    afterwards are first-touched on that NUMA node, so every rank streams its frames over its own PCIe root instead of
    across the socket interconnect.  Best effort: returns {'bus_id', 'numa_node', 'cpus'} or None when the topology cannot
    be read (no NVML, no sysfs entry, single-node machine reporting -1)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = int(visible.split(',')[device_index]) if visible and visible.split(',')[device_index].isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        sysfs = '/sys/bus/pci/devices/' + bus.lower()[-12:]
        node = int(open(sysfs + '/numa_node').read())
        cpus = parse_cpulist(open(sysfs + '/local_cpulist').read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if node < 0 or not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {'bus_id': bus, 'numa_node': node, 'cpus': len(allowed)}
    except Exception:       # pylint: disable=broad-except
        return None


def chunk_ranges(nframes: int, chunk_size: int, chunk_overlap: int = 0) -> List[range]:
    """Frame ranges of the chunks of a session, exactly the reference's gen_batch_sequence (io/util.py:24-35, offset 0):
    chunks of `chunk_size` frames starting every `chunk_size - chunk_overlap` frames -- the first `chunk_overlap` frames of
    every later chunk repeat the end of the previous one and are dropped by the writer (`data['offset']`)."""
    if chunk_size <= chunk_overlap:
        raise ValueError(f'chunk_size ({chunk_size}) must exceed chunk_overlap ({chunk_overlap})')
    seq = range(nframes)
    return [seq[i:i + chunk_size] for i in range(0, nframes - chunk_overlap, chunk_size - chunk_overlap)]


def shard_chunks(n_chunks: int, rank: int, world: int) -> range:
    """Contiguous chunk indices owned by `rank` out of `world` ranks."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f'bad rank/world {rank}/{world}')
    per = -(-n_chunks // world)
    return range(min(rank * per, n_chunks), min((rank + 1) * per, n_chunks))


def shard_frames(nframes: int, chunk_size: int, rank: int, world: int) -> range:
    """Frame range owned by `rank` (whole chunks only)."""
    chunks = chunk_ranges(nframes, chunk_size)
    mine = shard_chunks(len(chunks), rank, world)
    if len(mine) == 0:
        return range(0, 0)
    return range(chunks[mine.start].start, chunks[mine.stop - 1].stop)


class ShardedExtractor:
    """Runs `process_chunk(chunk_index, frame_range) -> dict of per-frame numpy arrays` over this rank's chunks.

    `process_chunk` is the GPU pipeline on a real run (ProduceFramesStep -> InferenceStep ->
    ProcessFeaturesStep on this rank's device); the CPU tests inject a stand-in so the partition / ordering /
    gather logic is exercised under gloo without a GPU."""

    def __init__(self, nframes: int, chunk_size: int, process_chunk: Callable[[int, range], Dict[str, np.ndarray]],
                 rank: int = 0, world: int = 1):
        self.nframes, self.chunk_size = int(nframes), int(chunk_size)
        self.process_chunk = process_chunk
        self.rank, self.world = int(rank), int(world)
        self.chunks = chunk_ranges(self.nframes, self.chunk_size)
        self.mine = shard_chunks(len(self.chunks), self.rank, self.world)

    def run(self) -> Dict[str, np.ndarray]:
        """Extract this rank's chunks; per-frame arrays are concatenated in frame order, plus `frame_idxs`."""
        parts: List[Dict[str, np.ndarray]] = []
        for ci in self.mine:
            res = dict(self.process_chunk(ci, self.chunks[ci]))
            res['frame_idxs'] = np.arange(self.chunks[ci].start, self.chunks[ci].stop)
            parts.append(res)
        return concat_results(parts)

    def gather(self, local: Dict[str, np.ndarray]) -> Optional[Dict[str, np.ndarray]]:
        """Control-plane gather of the per-frame tables to rank 0 (None elsewhere).  Not on the data path."""
        if self.world == 1:
            return local
        import torch.distributed as dist
        bucket = [None] * self.world if self.rank == 0 else None
        dist.gather_object(local, bucket, dst=0)
        if self.rank != 0:
            return None
        merged = concat_results([b for b in bucket if b])
        order = np.argsort(merged['frame_idxs'], kind='stable')
        return {k: v[order] for k, v in merged.items()}


def concat_results(parts: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    if not parts:
        return {}
    keys = parts[0].keys()
    return {k: np.concatenate([np.asarray(p[k]) for p in parts], axis=0) for k in keys}


# ---------------------------------------------------------------------------------------------------------------------
# the GPU runner: one process per GPU, the real step classes on this rank's chunks
# ---------------------------------------------------------------------------------------------------------------------
PER_FRAME_KEYS = ('depth_frames', 'mask_frames', 'flips', 'centroid', 'orientation', 'axis_length', 'num_instances')


def run_session_shard(session, config: dict, rank: int = 0, world: int = 1, inference: str = 'synthetic',
                      device_index: Optional[int] = None) -> Dict[str, np.ndarray]:
    """This rank's contiguous chunk range of `session` through ProduceFramesStep -> InferenceStep -> ProcessFeaturesStep (the
    step classes of ref pipeline/*.py) on one GPU, results as per-frame numpy arrays in frame order.

    Chunks follow the reference's sequence (ref: io/util.py:24-35, incl. `chunk_overlap`); like the reference's writer (ref:
    pipeline/write_results_step.py:54-73, io/result.py:105-130) the first `offset` frames of every chunk but the session's
    first are dropped, so the concatenation over ranks holds every frame exactly once and equals the single-rank result.
    `inference`: 'synthetic' (ground-truth instances of a synthetic session: BASELINE configs[1]) or 'model' (R-CNN:
    config['model'] = 'random' or a .ts path).  No collective: ranks never talk to each other here."""
    import torch
    from .pipeline import InferenceStep, Pipeline, PipelineStep, ProcessFeaturesStep, ProduceFramesStep, SyntheticInferenceStep
    if device_index is not None:
        torch.cuda.set_device(device_index)
    cfg = dict(config)
    cfg['chunk_shard'] = (int(rank), int(world))
    cfg.setdefault('results_to_host', True)
    parts: List[Dict[str, np.ndarray]] = []

    def host(x):
        return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)

    class Collect(PipelineStep):
        def process(self, data):
            off = int(data['offset'])
            feats = data['features']
            part = {'frame_idxs': np.asarray(data['frame_idxs'])[off:],
                    'depth_frames': host(data['depth_frames'])[off:], 'mask_frames': host(data['mask_frames'])[off:],
                    'flips': host(feats['flips'])[off:], 'centroid': host(feats['features']['centroid'])[off:],
                    'orientation': host(feats['features']['orientation'])[off:],
                    'axis_length': host(feats['features']['axis_length'])[off:], 'num_instances': np.asarray(feats['num_instances'])[off:]}
            for k, v in data['scalars'].items():
                part['scalars/' + k] = host(v)[off:]
            for k, v in data['keypoints'].items():
                part['keypoints/' + k] = host(v)[off:]
            parts.append(part)
            return data

    pipe = Pipeline()
    infer = SyntheticInferenceStep(cfg, 'infer') if inference == 'synthetic' else InferenceStep(cfg, 'infer')
    steps = [pipe.add_step(ProduceFramesStep(session, cfg, 'produce')), pipe.add_step(infer),
             pipe.add_step(ProcessFeaturesStep(cfg, 'features')), pipe.add_step(Collect(cfg, 'collect'))]
    for a, b in zip(steps[:-1], steps[1:]):
        pipe.link(a, b)
    pipe.run()
    return concat_results(parts)


def gather_shards(local: Dict[str, np.ndarray], rank: int, world: int) -> Optional[Dict[str, np.ndarray]]:
    """Control-plane gather of the per-frame tables to rank 0 in frame order (None elsewhere); not on the data path."""
    if world == 1:
        return local
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0)
    if rank != 0:
        return None
    merged = concat_results([b for b in bucket if b])
    order = np.argsort(merged['frame_idxs'], kind='stable')
    return {k: v[order] for k, v in merged.items()}
