from .pipeline import Pipeline, WorkerError  # noqa: F401
from .pipeline_step import PipelineStep, ProcessPipelineStep, ProducerPipelineStep, ThreadPipelineStep  # noqa: F401
from .produce_frames_step import ProduceFramesStep  # noqa: F401
from .inference_step import InferenceStep, SyntheticInferenceStep  # noqa: F401
from .process_features_step import ProcessFeaturesStep  # noqa: F401
from .write_results_step import ResultStore, ResultWriterStep  # noqa: F401
