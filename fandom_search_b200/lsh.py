"""LSH-emulation parity mode (FANDOM_SEARCH_MODE=lsh).

The reference indexes the script windows in `number_of_hashes` (15) nearpy
RandomBinaryProjections tables of `hash_dimensions` (14) random hyperplanes each
(search.py:112-116) and only compares a fan window with the windows that share one of its 15
bucket keys (search.py:178).  The hyperplanes are drawn un-seeded, so two reference runs
differ; to reproduce ONE run exactly the hyperplanes must be given.  This module draws them
the way a seeded nearpy does -- table i: numpy.random.RandomState(seed_i).randn(bits, dim)
with seed_i = (seed + crc32('rbp{i}')) mod 2**32, the convention of the oracle's nearpy
stand-in (oracle/shims/nearpy/hashes.py) -- and hands them to the device index, whose
lsh_first_table_kernel marks every exhaustive match with the first table in which the two
windows collide (or none).  The host then keeps only colliding pairs, ordered as nearpy orders
candidates (distance, first table, script position).
"""
import zlib

import numpy as np


def table_seed(seed, i):
    return (int(seed) + zlib.crc32(('rbp%d' % i).encode('utf-8'))) % (2 ** 32)


class LshEmulation:
    def __init__(self, number_of_hashes, hash_dimensions, vector_dim, seed):
        self.n_tables = int(number_of_hashes)
        self.n_bits = int(hash_dimensions)
        self.vector_dim = int(vector_dim)
        self.seed = int(seed)
        self.normals = np.concatenate(
            [np.random.RandomState(table_seed(seed, i)).randn(self.n_bits, self.vector_dim)
             for i in range(self.n_tables)], axis=0)

    def install(self, device_index):
        device_index.set_lsh(self.normals, self.n_tables, self.n_bits)
