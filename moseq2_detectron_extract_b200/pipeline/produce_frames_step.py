"""Producer step (ref: pipeline/produce_frames_step.py:11-48): iterate the session chunk by chunk and apply
`prep_raw_frames` -- here the fused CUDA prep kernel; the prepared chunk stays on the GPU."""
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
        def prep_on_device(frames, **kw):
            if isinstance(frames, torch.Tensor):
                out = prep_raw_frames(frames, **kw)          # CUDA tensor, or pinned host tensor consumed zero-copy
                return out if out.is_cuda else out.cuda()
            return prep_raw_frames(torch.from_numpy(frames).cuda(non_blocking=True), **kw)
        self.prep_frames = partial(prep_on_device, bground_im=self.session.bground_im, roi=self.session.roi,
                                   vmin=self.config['min_height'], vmax=self.config['max_height'])
        self.iterator = self.session.iterate(self.config['chunk_size'], self.config['chunk_overlap'])
        self.iterator.attach_filter(stream='depth', filterer=self.prep_frames)
        self.enumerator = enumerate(self.iterator)

    def process(self, data) -> Optional[dict]:
        try:
            i, (frame_idxs, raw_frames) = next(self.enumerator)
        except StopIteration:
            return None
        out = {'batch': i, 'chunk': raw_frames, 'frame_idxs': frame_idxs,
               'offset': self.config['chunk_overlap'] if i > 0 else 0}
        last = getattr(self.session, '_last', None)          # synthetic sessions expose their ground-truth instances
        if last is not None:
            out['synthetic_instances'] = last
        self.update_progress(int(raw_frames.shape[0]))
        return out
