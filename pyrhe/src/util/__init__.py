from pyrhe_b200.util import *  # noqa: F401,F403
from pyrhe_b200.util import Logger  # noqa: F401
