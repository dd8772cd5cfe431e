"""Producer step (ref: pipeline/produce_frames_step.py:11-48): iterate the session chunk by chunk and apply
`prep_raw_frames` -- here the fused CUDA prep kernel; the prepared chunk stays on the GPU."""
import collections
from functools import partial
from typing import Optional

import torch

from ..proc.proc import prep_raw_frames
from .pipeline_step import ProducerPipelineStep


class ProduceFramesStep(ProducerPipelineStep):
    def __init__(self, session, config: dict, name: Optional[str] = None, **kwargs) -> None:
        super().__init__(config, name, **kwargs)
        self.session = session

    def initialize(self):
        # pinned host chunks are read by the prep kernel straight from host memory (zero-copy): the buffer must outlive the
        # launch, and torch's host allocator knows nothing about a raw-pointer read -- every pinned chunk is held here with an
        # event recorded after its prep launch and released only once that event has completed (at most 2 in flight)
        self._inflight = collections.deque()

        def prep_on_device(frames, **kw):
            # the prep kernel's second output (one bit per positive pixel) rides along on the chunk tensor to ProcessFeaturesStep
            bits = []
            kw = dict(kw, positive_bits_out=bits)
            out = prep_any(frames, **kw)
            if bits and isinstance(out, torch.Tensor):
                out.msq_positive_bits = bits[0]
            return out

        def prep_any(frames, **kw):
            if isinstance(frames, torch.Tensor):
                out = prep_raw_frames(frames, **kw)          # CUDA tensor, or pinned host tensor consumed zero-copy
                if not frames.is_cuda and frames.is_pinned():
                    done = torch.cuda.Event()
                    done.record()
                    self._inflight.append((frames, done))
                    while len(self._inflight) > 2:
                        self._inflight.popleft()[1].synchronize()
                return out if out.is_cuda else out.cuda()
            return prep_raw_frames(torch.from_numpy(frames).cuda(non_blocking=True), **kw)
        self.prep_frames = partial(prep_on_device, bground_im=self.session.bground_im, roi=self.session.roi,
                                   vmin=self.config['min_height'], vmax=self.config['max_height'],
                                   fix_invalid_pixels=self.config.get('fix_invalid_pixels', True))
        self.iterator = self.session.iterate(self.config['chunk_size'], self.config['chunk_overlap'])
        self.iterator.attach_filter(stream='depth', filterer=self.prep_frames)
        # one process per GPU: config['chunk_shard'] = (rank, world) keeps this rank's contiguous range of the session's chunks
        # (shard.shard_chunks); chunk numbers and the overlap offset stay those of the whole session
        first = 0
        shard = self.config.get('chunk_shard')
        if shard is not None:
            from ..shard import shard_chunks
            mine = shard_chunks(len(self.iterator.batches), int(shard[0]), int(shard[1]))
            first = mine.start
            self.iterator.batches = self.iterator.batches[mine.start:mine.stop]
        self.enumerator = enumerate(self.iterator, start=first)

    def process(self, data) -> Optional[dict]:
        try:
            i, (frame_idxs, raw_frames) = next(self.enumerator)
        except StopIteration:
            return None
        out = {'batch': i, 'chunk': raw_frames, 'frame_idxs': frame_idxs,
               'offset': self.config['chunk_overlap'] if i > 0 else 0}
        bits = getattr(raw_frames, 'msq_positive_bits', None)
        if bits is not None:
            out['positive_bits'] = bits
        last = getattr(self.session, '_last', None)          # synthetic sessions expose their ground-truth instances
        if last is not None:
            out['synthetic_instances'] = last
        self.update_progress(int(raw_frames.shape[0]))
        return out

    def shutdown(self):
        while getattr(self, '_inflight', None):
            self._inflight.popleft()[1].synchronize()
