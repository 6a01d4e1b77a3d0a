from .file_processing import *  # noqa: F401,F403
from .types import *  # noqa: F401,F403
from .logger import Logger  # noqa: F401
