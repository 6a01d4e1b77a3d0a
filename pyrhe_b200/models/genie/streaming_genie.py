"""StreamingGENIE (/root/reference/pyrhe/src/models/genie/streaming_genie.py:10)."""
from ...base import StreamingBase
from .genie import GENIE


class StreamingGENIE(GENIE, StreamingBase):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
