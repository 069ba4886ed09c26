"""mavlm_b200 -- B200-native (sm_100a) visual-memory path of Memory-Augmented-VLM.

Directory name `memory-augmented-vlm_b200` is not a valid Python identifier; import it as
`mavlm_b200` (alias module at the repo root) or `importlib.import_module("memory-augmented-vlm_b200")`.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + C ABI -> libmavlm.so), `_lib.py`
(ctypes binding), `ops.py` (tensor-level wrappers), `modules.py` (mirror of the reference's module
interface), `pipeline.py` (chunk scheduler + fused path), `patch.py` (drop-in for a LLaVA model),
`dist.py` (video sharding over ranks), `legacy.py` (the reference's older memory builders and scene
segmentation).
"""
from . import _lib  # noqa: F401
from .modules import (Attention, Config, MemoryFuser, MemoryFuserMLP, Residual, TemporalPositionalEncoding, TransformerLayer,
                      TransformerProjector, VisionProjector, build_memory_fuser, build_vision_projector,
                      fine_frame_indices, get_2dPool, sample_frame_indices, uniform_segment_variant)
from .patch import convert_rmt, patch_llava
from . import preprocess  # noqa: F401
from . import legacy  # noqa: F401  (Flash-VStream-style memories, scene segmentation)
from .preprocess import SigLipImageProcessor, frames_preprocess
from .splice import IGNORE_INDEX, IMAGE_TOKEN_INDEX, splice_text_and_vision
from .pipeline import (FRAME_PROMPT_IDS, MEMORY_PROMPT_IDS, GraphedPipeline, GraphedTrainStep, HostStreamEncoder,
                       VisualMemoryPipeline)
from . import dist  # noqa: F401  (video sharding, frame-sharded pre-pass)

__all__ = [
    "Attention", "Config", "MemoryFuser", "MemoryFuserMLP", "Residual", "TemporalPositionalEncoding", "TransformerLayer",
    "TransformerProjector", "VisionProjector", "build_memory_fuser", "build_vision_projector", "fine_frame_indices",
    "get_2dPool", "sample_frame_indices", "uniform_segment_variant", "VisualMemoryPipeline", "GraphedPipeline", "GraphedTrainStep", "HostStreamEncoder", "MEMORY_PROMPT_IDS",
    "FRAME_PROMPT_IDS", "SigLipImageProcessor", "frames_preprocess", "patch_llava", "convert_rmt", "splice_text_and_vision", "IGNORE_INDEX", "IMAGE_TOKEN_INDEX",
]
