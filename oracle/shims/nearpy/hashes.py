"""[recalled] nearpy.hashes.RandomBinaryProjections (dense path only)."""
import os
import zlib

import numpy


def shim_seed(hash_name):
    """Seed used for `hash_name` when NEARPY_SHIM_SEED is set, else None (unseeded, as the
    reference constructs it at search.py:114-115)."""
    base = os.environ.get('NEARPY_SHIM_SEED')
    if base is None or base == '':
        return None
    return (int(base) + zlib.crc32(hash_name.encode('utf-8'))) % (2 ** 32)


class LSHash(object):
    def __init__(self, hash_name):
        self.hash_name = hash_name


class RandomBinaryProjections(LSHash):
    def __init__(self, hash_name, projection_count, rand_seed=None):
        super(RandomBinaryProjections, self).__init__(hash_name)
        self.projection_count = projection_count
        self.dim = None
        self.normals = None
        if rand_seed is None:
            rand_seed = shim_seed(hash_name)
        self.rand = numpy.random.RandomState(rand_seed)

    def reset(self, dim):
        if self.dim != dim:
            self.normals = None
        self.dim = dim
        if self.normals is None:
            self.normals = self.rand.randn(self.projection_count, dim)

    def hash_vector(self, v, querying=False):
        projection = numpy.dot(self.normals, v)
        return [''.join(['1' if x > 0.0 else '0' for x in projection])]
