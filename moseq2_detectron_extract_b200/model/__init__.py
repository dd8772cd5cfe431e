from .instances import Boxes, Instances  # noqa: F401
from .util import create_empty_instances, detector_postprocess, outputs_to_instances, paste_masks  # noqa: F401
