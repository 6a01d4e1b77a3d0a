"""Option enums of the public API.  Members and values are part of the drop-in contract
(/root/reference/pyrhe/src/util/types.py:3-9): the models compare the `.value` strings."""
from enum import Enum

GenoImputeMethod = Enum("GenoImputeMethod", {"BINARY": "binary", "MEAN": "mean"})
CovImputeMethod = Enum("CovImputeMethod", {"IGNORE": "ignore", "MEAN": "mean"})

__all__ = ["GenoImputeMethod", "CovImputeMethod"]
