"""[recalled] nearpy.storage.MemoryStorage."""


class MemoryStorage(object):
    def __init__(self):
        self.buckets = {}

    def store_vector(self, hash_name, bucket_key, v, data):
        self.buckets.setdefault(hash_name, {}).setdefault(bucket_key, []).append((v, data))

    def get_bucket(self, hash_name, bucket_key):
        return self.buckets.get(hash_name, {}).get(bucket_key, [])

    def clean_all_buckets(self):
        self.buckets = {}
