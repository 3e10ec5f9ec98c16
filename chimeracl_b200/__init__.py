"""chimera-b200: B200-native implementation of chimeraCL's per-step PIC hot path
behind chimeraCL's own Python wrapper API (Particles / Grid / Solver / Transformer /
PIC_loop, dict-of-arrays Args/DataDev containers, methods/ mixins)."""
__version__ = "0.1.0"
