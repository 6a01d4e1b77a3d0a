from pyrhe_b200.models import *  # noqa: F401,F403
