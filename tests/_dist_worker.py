"""Worker of tests/test_parallel_gloo.py: one rank of a world_size-N `analyze` run on CPU
(gloo), with the device replaced by the numpy checker index."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fandom_search_b200.engine as engine_mod  # noqa: E402
from fandom_search_b200 import search  # noqa: E402
from fandom_search_b200.lexicon import Lexicon, py_hash_seed0  # noqa: E402
from tests.numpy_index import NumpyIndex  # noqa: E402


def main():
    golden, workdir = sys.argv[1], sys.argv[2]
    if os.environ.get('FS_TEST_REAL_DEVICE') != '1':
        engine_mod.DeviceIndex = NumpyIndex
    search.set_pipeline(search.Pipeline(
        Lexicon.from_npz(os.path.join(golden, "lexicon.npz"), hash_fn=py_hash_seed0)))
    listing = open(os.path.join(golden, "listing.txt")).read().split()
    real_listdir = os.listdir
    os.listdir = lambda d: list(listing) if str(d) == "fanworks" else real_listdir(d)
    os.chdir(workdir)
    args = argparse.Namespace(fan_works="fanworks", script="script.txt", skip_works=-1, num_works=-1)
    if os.environ.get('FS_TEST_FAIL_RANK') == os.environ.get('RANK'):
        def broken(self, filenames):
            raise OSError("simulated unreadable fanwork on rank %s" % os.environ.get('RANK'))
        search.AnnIndexSearch.prepare = broken
    search.analyze(args, chunk_size=16)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
