"""Per-axis metrics of the MLFriends constrainer (reference clustering/sdml.py:25-88).

Host-side numpy, O(n * ndim) per region rebuild: not device work, but the constrainer mirror
(massivedatans_b200/hiermetriclearn.py) needs them, and the members handed to the neighbour
kernels must be the same doubles the reference would hand to cneighbors.c -- so every
operation below is the reference's, in the reference's order (mean, centred std, power-of-two
rounding; sdml.py:68-77).
"""
import numpy


class _Metric(object):
    def __eq__(self, other):
        # sdml.py:36-37 compares attribute dictionaries; arrays are compared by value here so
        # that the comparison has a truth value under numpy >= 1.25 as well
        a, b = self.__dict__, other.__dict__
        if a.keys() != b.keys():
            return False
        return all(numpy.array_equal(a[k], b[k]) for k in a)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None


class IdentityMetric(_Metric):
    """sdml.py:25-37: the unit metric."""

    def fit(self, x):
        pass

    def transform(self, x):
        return x

    def untransform(self, y):
        return y


class SimpleScaling(_Metric):
    """sdml.py:39-60: subtract the mean, divide by the per-axis standard deviation."""

    def __init__(self, verbose=False):
        self.verbose = verbose

    def _scale_of(self, centred):
        return numpy.std(centred, axis=0)

    def fit(self, X, W=None):
        self.mean = numpy.mean(X, axis=0)
        self.scale = self._scale_of(X - self.mean)

    def transform(self, x):
        return (x - self.mean) / self.scale

    def untransform(self, y):
        return y * self.scale + self.mean


class TruncatedScaling(SimpleScaling):
    """sdml.py:62-88: like SimpleScaling, with every axis scale rounded to a power of two
    relative to 1.001 x the largest one, so that the metric does not random-walk."""

    def _scale_of(self, centred):
        std = numpy.std(centred, axis=0)
        top = std.max() * 1.001
        steps = (-numpy.log2(std / top)).astype(int)
        return 2 ** (steps.astype(float))
