"""[recalled] nearpy.utils.utils.unitvec for dense vectors."""
import numpy


def unitvec(vec):
    vec_norm = numpy.linalg.norm(vec)
    if vec_norm > 0.0:
        return vec / vec_norm
    return vec
