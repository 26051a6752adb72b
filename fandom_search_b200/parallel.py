"""Work-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards by cluster of fanworks: cluster i is searched entirely by rank
i % WORLD_SIZE against a replicated script index (the reference's own unit of parallelism is
the file, search.py:382-385).  There is NO collective on the data path: every rank writes the
batch CSVs of its clusters, and after one barrier rank 0 assembles the aggregate CSV
(search.py:388-399) from those files.  (analyze_scripts, the multi-script pass, still returns its
record lists to rank 0 with gather_object.)
"""
import os


def init_process_group():
    """Initialise torch.distributed from the torchrun environment if needed.
    Returns (rank, world).  nccl when CUDA is available, gloo otherwise (CPU tests)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1') or 1)
    rank = int(os.environ.get('RANK', '0') or 0)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29531')
        if torch.cuda.is_available():
            local = int(os.environ.get('LOCAL_RANK', str(rank)) or 0)
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', rank=rank, world_size=world,
                                    device_id=torch.device('cuda', local))
        else:
            dist.init_process_group('gloo', rank=rank, world_size=world)
    return rank, world


def barrier():
    """All ranks wait for each other (host-side control only; nothing on the data path)."""
    import torch
    import torch.distributed as dist
    rank, world = init_process_group()
    if world > 1:
        if torch.cuda.is_available():
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def cluster_owner(cluster_index, world):
    return cluster_index % world


def gather_cluster_records(my_records, rank, world):
    """my_records: {cluster index: record list} of this rank -> on rank 0 the union over
    ranks (cluster indices are disjoint by construction); other ranks get {}."""
    import torch.distributed as dist
    init_process_group()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(my_records, gathered, dst=0)
    if rank != 0:
        return {}
    merged = {}
    for part in gathered:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError("cluster %d searched by two ranks" % k)
            merged[k] = v
    return merged
