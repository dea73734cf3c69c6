"""contrastboundary_b200 — B200-native (sm_100a) point-cloud operator stack behind the operator
API of LiyaoTang/contrastBoundary.  See DESIGN.md.  No CPU fallback: operators raise if
libcbops.so (python -m contrastboundary_b200.build) is missing."""
__version__ = "0.1.0"
