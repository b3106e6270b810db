"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol
include/mdns_b200.h declares, its host helpers are exact, and compute entry points fail
loudly (no CPU fallback) when no device is present."""
import ctypes
import math
import os
import re
import subprocess

import numpy
import pytest

from conftest import ROOT
from massivedatans_b200 import _lib


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'mdns_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mdns_[a-z0-9_]+)\s*\(', text)))


def test_header_and_bindings_agree():
    names = _header_functions()
    assert len(names) >= 35
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if ' T ' in l)
    for name in _header_functions():
        assert name in exported, name
        assert getattr(lib, name) is not None
    assert lib.mdns_version() >= 100


def test_dropin_veneers_export_reference_symbols():
    want = {'clike.so': {'like'}, 'cmuselike.so': {'like'},
            'cneighbors.so': {'most_distant_nearest_neighbor', 'is_within_distance_of',
                              'count_within_distance_of', 'bootstrapped_maxdistance'}}
    for so, syms in want.items():
        path = os.path.join(_lib.DROPIN_DIR, so)
        out = subprocess.check_output(['nm', '-D', '--defined-only', path]).decode()
        exported = set(l.split()[-1] for l in out.splitlines() if ' T ' in l)
        assert syms <= exported, (so, exported)
        ctypes.CDLL(path)       # resolves against libmdns_b200.so through its rpath


def test_sqrt_threshold_is_exact():
    lib = _lib.load()
    rs = numpy.random.RandomState(0)
    radii = numpy.concatenate([rs.uniform(0, 2, 200), 10 ** rs.uniform(-300, 300, 200),
                               [0.1, 0.25, 1.0, 3.0, 1e-170, 1e-160, 5e-324, 1.3e154, 1.4e154,
                                1.7e308]])
    for r in radii:
        T = lib.mdns_sqrt_threshold(float(r))
        cand = [T]
        lo = hi = T
        for _ in range(4):
            lo = numpy.nextafter(lo, -numpy.inf)
            hi = numpy.nextafter(hi, numpy.inf)
            cand += [lo, hi]
        cand += [float(r) * float(r), 0.0]
        for d in cand:
            if not (d >= 0) or math.isinf(d):
                continue
            assert (math.sqrt(d) < r) == (d < T), (r, d, T)
    assert lib.mdns_sqrt_threshold(0.0) == 0.0
    assert lib.mdns_sqrt_threshold(-1.0) == 0.0
    assert lib.mdns_sqrt_threshold(float('nan')) == 0.0
    assert lib.mdns_sqrt_threshold(float('inf')) == float('inf')


@pytest.mark.skipif(_lib.load().mdns_device_count() > 0, reason='a GPU is present')
def test_no_cpu_fallback():
    from massivedatans_b200.clustering import neighbors
    from massivedatans_b200.likelihood import ResidentDataset, make_multi_loglikelihood
    with pytest.raises(_lib.MdnsError):
        ResidentDataset(numpy.zeros(4), numpy.zeros((4, 3)))
    with pytest.raises(_lib.MdnsError):
        make_multi_loglikelihood(numpy.zeros(4), numpy.zeros((4, 3)))
    with pytest.raises(_lib.MdnsError):
        neighbors.count_within_distance_of(numpy.zeros((3, 2)), 0.1, numpy.zeros((2, 2)))
    h = ctypes.c_void_p()
    y = numpy.zeros((4, 3))
    rc = _lib.load().mdns_dataset_create(None, y.ctypes.data, None, 3, 4, None, 0, ctypes.byref(h))
    assert rc == -2 and b'no CUDA device' in _lib.load().mdns_last_error()
