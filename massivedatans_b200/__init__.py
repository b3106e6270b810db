"""B200-native hot path of massivedatans (collaborative nested sampling).

Batched Gaussian chi-square log-likelihood (reference: clike.c, cmuselike.c) and
the RadFriends neighbour tests (reference: clustering/cneighbors.c), as
hand-written CUDA for sm_100a behind a ctypes-loaded C-ABI shared library
(``libmdns_b200.so``, see ``include/mdns_b200.h``).  There is no CPU fallback:
importing the compute entry points without the built library raises.
"""
__version__ = '0.1.0'
