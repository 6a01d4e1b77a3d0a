from pyrhe_b200.base import Base, StreamingBase  # noqa: F401
