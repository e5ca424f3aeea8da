"""B200-native hot path of agussomacal/ROMHighContrast.

`romhighcontrast_b200.lib` mirrors the reference's `src/lib` API (SolutionsManagers, ReducedBasis,
Estimators); `romhighcontrast_b200.engine.Engine` is the thin device layer over the C ABI
(`include/romhc.h`, `libromhc.so`).  There is no CPU fallback.
"""
__version__ = "0.1.0"
