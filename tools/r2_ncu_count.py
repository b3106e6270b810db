#!/usr/bin/env python
"""A few full-count calls at n members x m candidates (for ncu): n m [ndim]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.clustering import neighbors  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
ndim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
xx, yy = synth.members_and_candidates(n, m, ndim)
r = 0.5 * n ** (-1.0 / ndim)
for _ in range(3):
    c = neighbors.count_within_distance_of(xx, r, yy)
print(int(c.sum()))
