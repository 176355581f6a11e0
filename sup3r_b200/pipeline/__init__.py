"""sup3r_b200.pipeline: forward-pass chunk tiler (``sup3r.pipeline`` API surface)."""
from .forward_pass import ForwardPass
from .slicer import ForwardPassSlicer
from .strategy import ArrayInputHandler, ForwardPassChunk, ForwardPassStrategy

__all__ = ["ForwardPass", "ForwardPassSlicer", "ForwardPassStrategy", "ForwardPassChunk",
           "ArrayInputHandler"]
