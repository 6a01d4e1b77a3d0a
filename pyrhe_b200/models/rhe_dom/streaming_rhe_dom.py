"""StreamingRHE_DOM (/root/reference/pyrhe/src/models/rhe_dom/streaming_rhe_dom.py:7)."""
from ...base import StreamingBase
from .rhe_dom import RHE_DOM


class StreamingRHE_DOM(RHE_DOM, StreamingBase):
    pass
