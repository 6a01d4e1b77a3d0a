from .rhe import RHE  # noqa: F401
from .streaming_rhe import StreamingRHE  # noqa: F401
