"""Drop-in spelling of the reference package (`src.lib` / `lib`, both are used by its callers:
/root/reference/src/experiments/HighContrast.py:15-17,26).  Re-exports romhighcontrast_b200.lib."""
