"""wgdi_b200 — B200-native drop-in for the `wgdi -icl` (improved collinearity) hot path.

Host side (this package) mirrors the reference's interface for the path:
`run_colliearity.mycollinearity(options).run()` and `collinearity.collinearity(options, points).run()`;
the work runs in hand-written sm_100a CUDA kernels behind the C ABI of include/icl_b200.h.
"""
__version__ = "0.1.0"
