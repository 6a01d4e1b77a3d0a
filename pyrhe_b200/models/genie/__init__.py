from .genie import GENIE  # noqa: F401
from .streaming_genie import StreamingGENIE  # noqa: F401
