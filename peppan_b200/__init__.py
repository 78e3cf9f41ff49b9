"""peppan_b200 -- B200-native similarity search and clustering for PEPPAN's hot path.

Host-side mirror of the reference's modules/uberBlast.py and modules/clust.py entry points on top
of libpeppan_b200.so (CUDA, sm_100a) reached through ctypes on numpy buffers.  There is no CPU
fallback: every compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = '0.1.0'
