"""StreamingRHE (/root/reference/pyrhe/src/models/rhe/streaming_rhe.py:6)."""
from ...base import StreamingBase
from .rhe import RHE


class StreamingRHE(RHE, StreamingBase):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
