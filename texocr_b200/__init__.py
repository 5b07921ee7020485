"""texocr_b200 -- B200-native inference path of TeXOCR behind the reference's OCRModel API."""
import os as _os

# Decode branches and batches in flight each run on their own stream; the default of 8 hardware work queues makes streams
# share queues (false serialisation).  Only effective when set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import spec, synth  # noqa: F401
from .model import OCRModel, create_model  # noqa: F401
from ._lib import Engine, load_library  # noqa: F401
from . import checkpoint, detok, preprocess  # noqa: F401
from .wrapper import TeXOCRWrapper  # noqa: F401
from .pipeline import GeneratePipeline  # noqa: F401

__all__ = ["OCRModel", "create_model", "Engine", "load_library", "TeXOCRWrapper", "GeneratePipeline", "spec", "synth", "checkpoint", "detok", "preprocess"]
