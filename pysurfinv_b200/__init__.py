"""pysurfinv_b200 -- B200-native batched surface-wave dispersion forward solver (drop-in for the
``fast_surf.fast_surf`` hot path of 001cat/pySurfInv).  Host side is Python over a C-ABI CUDA library."""
