from .rhe_dom import RHE_DOM  # noqa: F401
from .streaming_rhe_dom import StreamingRHE_DOM  # noqa: F401
