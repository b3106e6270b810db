"""B200-native hot path of massivedatans (collaborative nested sampling).

Batched Gaussian chi-square log-likelihood (reference: clike.c, cmuselike.c) and
the RadFriends neighbour tests (reference: clustering/cneighbors.c), as
hand-written CUDA for sm_100a behind a ctypes-loaded C-ABI shared library
(``libmdns_b200.so``, see ``include/mdns_b200.h``).  There is no CPU fallback:
importing the compute entry points without the built library raises.

Modules (each mirrors the reference file of the same role, same names and signatures):
  likelihood                   sample.py:78-108, musefuse.py:222-284,503-542 (callables, MUSE model)
  clustering.neighbors         clustering/neighbors.py
  clustering.radfriendsregion  clustering/radfriendsregion.py (+ device-side candidate generation)
  clustering.sdml              clustering/sdml.py
  hiermetriclearn              hiermetriclearn.py (MLFriends constrained draw, speculative batches)
  livepoints                   multi_nested_sampler.py:134-137,204-355,438-447,520-524 (live table)
  sharding                     data-set ranges over GPUs, first-accept count exchange
  synth                        seeded restatements of the reference's data generators
"""
__version__ = '0.1.0'
