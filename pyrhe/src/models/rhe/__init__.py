from pyrhe_b200.models.rhe import *  # noqa: F401,F403
from pyrhe_b200.models.rhe import RHE, StreamingRHE  # noqa: F401
