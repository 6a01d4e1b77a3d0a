"""Drop-in import paths of the reference package: `pyrhe.models` (README) and
`pyrhe.src.{base,models,util}` (run_rhe.py:4-7), all backed by pyrhe_b200."""
