"""Step runtime with the reference's contract (ref: pipeline/pipeline_step.py:12-192).

Same public surface (`initialize / process(data) -> data|None / finalize`, `update_progress`,
`write_message`, None as end-of-stream sentinel, one in-queue and any number of out-queues), but all steps
of one pipeline live in ONE process per GPU and run as threads: the data dicts carry CUDA tensors from step
to step, so nothing is pickled and full-frame masks never cross PCIe (the reference's InferenceStep moves
every Instances to the CPU, ref: pipeline/inference_step.py:68).
"""
from __future__ import annotations

import logging
import queue
import threading
import traceback
from typing import List, Optional, Union


class PipelineStep(threading.Thread):
    """One step of a Pipeline: pulls dicts from `in_queue`, pushes results to every queue in `out_queue`."""

    def __init__(self, config: dict, name: Optional[str] = None, **kwargs) -> None:
        super().__init__(name=name, daemon=True)
        self.step_name = name
        self.is_producer = False
        self.shutdown_event: Optional[threading.Event] = None
        self.progress: Optional[queue.Queue] = None
        self.in_queue: Optional[queue.Queue] = None
        self.out_queue: List[queue.Queue] = []
        self.is_complete = threading.Event()
        self.config = config
        self.error: Optional[str] = None

    # ---- progress / messages (ref: pipeline_step.py:28-70) ---------------------------------------
    def attach_progress(self, progress_queue: queue.Queue) -> None:
        self.progress = progress_queue

    def _post(self, item: dict) -> None:
        if self.progress is not None:
            self.progress.put(item)

    def reset_progress(self, total: int) -> None:
        self._post({'total': total})

    def update_progress(self, incremental_progress: int = 1) -> None:
        self._post({'update': incremental_progress})

    def write_message(self, message: str, level=logging.INFO, raise_exc: bool = False) -> None:
        self._post({'message': message, 'level': level, 'raise': raise_exc})

    def flush_progress(self) -> None:
        self._post({'flush': True})

    # ---- queues (ref: pipeline_step.py:72-96) ----------------------------------------------------
    def set_outputs(self, data) -> None:
        """Blocks while a consumer's queue is full, but gives up when the pipeline is shutting down (a consumer that died can no
        longer drain its queue: the producer must not hang on it holding CUDA tensors)."""
        for q in self.out_queue:
            while True:
                try:
                    q.put(data, timeout=0.1)
                    break
                except queue.Full:
                    if self.shutdown_event is not None and self.shutdown_event.is_set():
                        return

    def is_output_empty(self) -> bool:
        return all(q.empty() for q in self.out_queue)

    def signal_shutdown(self) -> None:
        if self.shutdown_event is not None:
            self.shutdown_event.set()

    def shutdown(self) -> None:
        """Called once when the step's loop ends (normally or not)."""

    @property
    def total_items(self) -> int:
        return self.config['nframes']

    # ---- main loop (ref: pipeline_step.py:106-160) -------------------------------------------------
    def run(self) -> None:
        try:
            self.reset_progress(self.total_items)
            assert self.shutdown_event is not None
            self.initialize()
            while not self.shutdown_event.is_set():
                if self.is_producer:
                    data = None
                else:
                    assert self.in_queue is not None
                    try:
                        data = self.in_queue.get(block=True, timeout=0.1)
                    except queue.Empty:
                        continue
                    if data is None:
                        self.set_outputs(None)
                        self.is_complete.set()
                        break
                out: Union[dict, None] = self.process(data)
                self.set_outputs(out)
                self.flush_progress()
                if self.is_producer and out is None:
                    self.is_complete.set()
                    break
        except Exception:  # pylint: disable=broad-except
            self.error = traceback.format_exc()
            self.write_message(self.error, level=logging.CRITICAL, raise_exc=True)
            self.finalize()
            self.signal_shutdown()
        finally:
            self.flush_progress()
            try:
                self.shutdown()          # end-of-stream hook (the reference leaves it commented out, pipeline_step.py:157):
            except Exception:            # writers close their files here; pylint: disable=broad-except
                self.error = (self.error or '') + traceback.format_exc()
            self.is_complete.set()

    def initialize(self) -> None:
        """Called once before the first batch."""

    def process(self, data: dict) -> Union[dict, None]:
        """Called for every batch."""

    def finalize(self) -> None:
        """Called on shutdown after an error."""


class ThreadPipelineStep(PipelineStep):
    """Step that runs in a thread (all of them do here)."""


class ProcessPipelineStep(PipelineStep):
    """Name kept for drop-in subclasses; runs as a thread of the per-GPU process."""


class ProducerPipelineStep(ThreadPipelineStep):
    """Step that only produces (ref: pipeline_step.py:185-192)."""

    def __init__(self, config: dict, name: Optional[str] = None, **kwargs) -> None:
        super().__init__(config, name, **kwargs)
        self.is_producer = True
