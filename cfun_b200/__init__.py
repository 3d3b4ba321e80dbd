"""cfun_b200 -- B200-native (sm_100a) implementation of the CFUN volumetric hot path.

Importing the package never touches the GPU; cfun_b200.ops / model load libcfun_b200.so (built in-tree by
cfun_b200.build) and raise if it is missing -- there is no CPU fallback."""
__version__ = "0.1.0"
