"""pyrhe_b200: B200-native (sm_100a) implementation of PyRHE's randomized Haseman-Elston
trace-estimation hot path, behind the reference's `pyrhe.models` API."""
__version__ = "0.1.0"
