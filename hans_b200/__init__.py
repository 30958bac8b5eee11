"""hans_b200 -- B200-native nested-sampling Monte Carlo walker engine (drop-in for the hs_alkane
hot path of omaradesida/hans).  See DESIGN.md.

Sub-modules import the CUDA extension (libhans_b200.so) and fail loudly when it is missing; the
pure host-side helpers (input parsing, file formats, Philox on the host) live in modules that do
not need it.
"""
__version__ = "0.1.0"
