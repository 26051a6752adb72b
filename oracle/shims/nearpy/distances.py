"""[recalled] nearpy.distances (dense path only)."""
import os

import numpy


class Distance(object):
    def distance(self, x, y):
        raise NotImplementedError


class CosineDistance(Distance):
    """1 - cos(angle(x, y)); the engine passes unit vectors on both sides."""

    def distance(self, x, y):
        if os.environ.get('NEARPY_SHIM_COSINE_RENORM', '0') == '1':
            return 1.0 - numpy.dot(x, y) / (numpy.linalg.norm(x) * numpy.linalg.norm(y))
        return 1.0 - numpy.dot(x, y)


class EuclideanDistance(Distance):
    def distance(self, x, y):
        return numpy.linalg.norm(x - y)
