"""Pipeline container (ref: pipeline/pipeline.py:12-136): add steps, link them, start, poll, shut down."""
from __future__ import annotations

import queue
import threading
import time
from typing import List

from .pipeline_step import PipelineStep


class WorkerError(RuntimeError):
    """A step raised; carries the formatted traceback(s) (ref: pipeline/progress.py:128-131)."""


class Pipeline:
    def __init__(self, queue_depth: int = 2) -> None:
        self.steps: List[PipelineStep] = []
        self.shutdown_event = threading.Event()
        self.progress: queue.Queue = queue.Queue()
        self.queue_depth = queue_depth

    def add_step(self, step: PipelineStep) -> PipelineStep:
        step.shutdown_event = self.shutdown_event
        step.attach_progress(self.progress)
        self.steps.append(step)
        return step

    def link(self, producer: PipelineStep, *consumers: PipelineStep) -> None:
        for consumer in consumers:
            q: queue.Queue = queue.Queue(maxsize=self.queue_depth)
            producer.out_queue.append(q)
            consumer.in_queue = q

    def start(self) -> None:
        for step in self.steps:
            step.start()

    def is_running(self) -> bool:
        return any(not s.is_complete.is_set() for s in self.steps) and not self.shutdown_event.is_set()

    def shutdown(self, wait_seconds: float = 3.0) -> None:
        self.shutdown_event.set()
        deadline = time.time() + wait_seconds
        for step in self.steps:
            step.join(timeout=max(0.0, deadline - time.time()))
        errors = [s.error for s in self.steps if s.error]
        if errors:
            raise WorkerError('\n'.join(errors))

    def run(self, poll: float = 0.01) -> None:
        self.start()
        while self.is_running():
            time.sleep(poll)
        self.shutdown()
