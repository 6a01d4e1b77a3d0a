from .base import Base  # noqa: F401
from .base_streaming import StreamingBase  # noqa: F401
