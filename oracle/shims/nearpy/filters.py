"""[recalled] nearpy.filters.NearestFilter / UniqueFilter."""


class VectorFilter(object):
    def filter_vectors(self, input_list):
        raise NotImplementedError


class NearestFilter(VectorFilter):
    def __init__(self, N):
        self.N = N

    def filter_vectors(self, input_list):
        try:
            sorted_list = sorted(input_list, key=lambda x: x[2])  # stable
            return sorted_list[:self.N]
        except Exception:
            return input_list


class UniqueFilter(VectorFilter):
    def filter_vectors(self, input_list):
        unique_dict = {}
        for v in input_list:
            unique_dict[v[1]] = v  # first position kept, later duplicate overwrites the value
        return list(unique_dict.values())
