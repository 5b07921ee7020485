"""texocr_b200 -- B200-native inference path of TeXOCR behind the reference's OCRModel API."""
from . import spec, synth  # noqa: F401
from .model import OCRModel, create_model  # noqa: F401
from ._lib import Engine, load_library  # noqa: F401
from . import checkpoint, detok, preprocess  # noqa: F401
from .wrapper import TeXOCRWrapper  # noqa: F401

__all__ = ["OCRModel", "create_model", "Engine", "load_library", "TeXOCRWrapper", "spec", "synth", "checkpoint", "detok", "preprocess"]
