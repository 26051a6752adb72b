"""TEST INFRASTRUCTURE -- stand-in for the `nearpy` package (oracle only, never shipped).

The reference imports nearpy un-pinned (/root/reference requirements.txt:2; the latest PyPI
release is NearPy 1.0.0) and it is not installed here, so this module RESTATES the published
behaviour of the handful of classes search.py touches (call sites search.py:114-123,178).
Parity against the real package is therefore *unpinned*: every statement below marked
[recalled] comes from knowledge of NearPy 1.0.0, not from source on this machine.

Extra knobs (environment variables; all default to the faithful behaviour):
  NEARPY_SHIM_SEED        int  -> RandomBinaryProjections(name) is seeded with
                                  (seed + crc32(name)) mod 2**32 instead of unseeded
  NEARPY_SHIM_EXHAUSTIVE  1    -> every query sees every stored vector (no LSH loss)
  NEARPY_SHIM_UNIQUE      0    -> neighbours() does not apply UniqueFilter by default
  NEARPY_SHIM_COSINE_RENORM 1  -> CosineDistance divides by |x||y| again
"""
import os
import zlib

import numpy

from . import hashes, distances, filters, storage  # noqa: F401  (attribute access as nearpy.hashes...)
from .hashes import RandomBinaryProjections
from .distances import CosineDistance, EuclideanDistance
from .filters import NearestFilter, UniqueFilter
from .storage import MemoryStorage
from .utils import unitvec

__version__ = "1.0.0-shim"


class Engine(object):
    """[recalled] nearpy.engine.Engine."""

    def __init__(self, dim, lshashes=None, distance=None, fetch_vector_filters=None,
                 vector_filters=None, storage=None):
        if lshashes is None:
            lshashes = [RandomBinaryProjections('default', 10)]
        self.lshashes = lshashes
        if distance is None:
            distance = EuclideanDistance()
        self.distance = distance
        if vector_filters is None:
            vector_filters = [NearestFilter(10)]
        self.vector_filters = vector_filters
        if fetch_vector_filters is None:
            fetch_vector_filters = [UniqueFilter()]
        self.fetch_vector_filters = fetch_vector_filters
        if storage is None:
            storage = MemoryStorage()
        self.storage = storage
        self.dim = dim
        for lshash in self.lshashes:
            lshash.reset(dim)
        self._exhaustive = os.environ.get('NEARPY_SHIM_EXHAUSTIVE', '0') == '1'
        self._apply_unique = os.environ.get('NEARPY_SHIM_UNIQUE', '1') != '0'
        self._all = []

    def store_vector(self, v, data=None):
        nv = unitvec(v)
        for lshash in self.lshashes:
            for bucket_key in lshash.hash_vector(v):
                self.storage.store_vector(lshash.hash_name, bucket_key, nv, data)
        if self._exhaustive:
            self._all.append((nv, data))

    def neighbours(self, v, distance=None, fetch_vector_filters=None, vector_filters=None):
        candidates = self._get_candidates(v)
        if fetch_vector_filters is None and self._apply_unique:
            fetch_vector_filters = self.fetch_vector_filters
        if fetch_vector_filters:
            candidates = self._apply_filter(fetch_vector_filters, candidates)
        if distance is None:
            distance = self.distance
        candidates = self._append_distances(v, distance, candidates)
        if not vector_filters:
            vector_filters = self.vector_filters
        candidates = self._apply_filter(vector_filters, candidates)
        return candidates

    def _get_candidates(self, v):
        if self._exhaustive:
            return list(self._all)
        candidates = []
        for lshash in self.lshashes:
            for bucket_key in lshash.hash_vector(v, querying=True):
                candidates.extend(self.storage.get_bucket(lshash.hash_name, bucket_key))
        return candidates

    def _apply_filter(self, filters_, candidates):
        if filters_:
            for f in filters_:
                candidates = f.filter_vectors(candidates)
        return candidates

    def _append_distances(self, v, distance, candidates):
        if distance:
            nv = unitvec(v)
            candidates = [(x[0], x[1], distance.distance(x[0], nv)) for x in candidates]
        return candidates

    def candidate_count(self, v):
        return len(self._get_candidates(v))

    def clean_all_buckets(self):
        self.storage.clean_all_buckets()
        self._all = []
