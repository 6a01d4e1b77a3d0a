from .rhe import RHE, StreamingRHE  # noqa: F401
from .rhe_dom import RHE_DOM, StreamingRHE_DOM  # noqa: F401
from .genie import GENIE, StreamingGENIE  # noqa: F401
