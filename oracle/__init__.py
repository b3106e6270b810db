"""CPU oracle for the massivedatans hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``massivedatans_b200/`` imports this package.  Allowed users:
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.

Three layers, all on the host CPU:

* ``oracle.ref``  -- the UNMODIFIED reference C (``clike.c``, ``cmuselike.c``,
  ``clustering/cneighbors.c``) compiled by ``oracle/Makefile`` into
  ``oracle/_ref/*.so`` and bound with the reference's own ctypes argtypes
  (``sample.py:85-96``, ``musefuse.py:509-517``, ``clustering/neighbors.py:100-167``).
* ``oracle.port`` -- our plain-C restatement ``oracle/mdns_oracle.c``
  (``liboracle.so``), bit-identical to ``oracle.ref`` (pinned by
  ``tests/test_oracle_vs_reference.py`` and the fixtures in ``tests/golden``).
* ``oracle.np``   -- numpy/scipy restatements of the same formulas
  (``sample.py:64-71``, ``musefuse.py:464-481``, ``clustering/neighbors.py:79-85``).
"""
from . import np_oracle as np  # noqa: F401
from . import port, ref  # noqa: F401
