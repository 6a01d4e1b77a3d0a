from pyrhe_b200.models.rhe_dom import RHE_DOM, StreamingRHE_DOM  # noqa: F401
