"""Option enums (same names/values as /root/reference/pyrhe/src/util/types.py:3-9)."""
from enum import Enum


class GenoImputeMethod(Enum):
    BINARY = "binary"
    MEAN = "mean"


class CovImputeMethod(Enum):
    IGNORE = "ignore"
    MEAN = "mean"
