from .file_processing import *  # noqa: F401,F403
from .types import *  # noqa: F401,F403
from .logger import Logger  # noqa: F401
from .mat_mul import to_tensor, mat_mul, elem_mul  # noqa: F401
