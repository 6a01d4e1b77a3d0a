from pyrhe_b200.models.genie import GENIE, StreamingGENIE  # noqa: F401
