"""Work-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards by cluster of fanworks: cluster i is searched entirely by rank
i % WORLD_SIZE against a replicated script index (the reference's own unit of parallelism is
the file, search.py:382-385).  There is NO collective on the data path: every rank writes the
batch CSVs of its clusters, and after one barrier rank 0 assembles the aggregate CSV
(search.py:388-399) from those files.  (analyze_scripts, the multi-script pass, still returns its
record lists to rank 0 with gather_object.)
"""
import os


def device_ordinal():
    """The CUDA device of this process -- the ONE place that decides it: FANDOM_SEARCH_DEVICE if
    set, else LOCAL_RANK (torchrun), else 0."""
    for key in ('FANDOM_SEARCH_DEVICE', 'LOCAL_RANK'):
        v = os.environ.get(key)
        if v not in (None, ''):
            return int(v)
    return 0


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
            local = device_ordinal()
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


def assign_clusters(sizes, world, policy=None):
    """Owner rank of every cluster.  policy 'balanced' (default under FANDOM_SEARCH_BALANCE unset):
    longest-processing-time greedy on the cluster sizes (bytes of text ~ tokens ~ windows), so that a
    corpus whose cluster count is not a multiple of the world size, or whose clusters differ in size,
    does not leave a tail on one GPU; 'roundrobin': cluster i -> rank i % world.
    Deterministic: every rank computes the same table from the same sizes."""
    if policy is None:
        policy = os.environ.get('FANDOM_SEARCH_BALANCE', 'balanced')
    n = len(sizes)
    if world <= 1:
        return [0] * n
    if policy == 'roundrobin':
        return [cluster_owner(i, world) for i in range(n)]
    if policy != 'balanced':
        raise ValueError("FANDOM_SEARCH_BALANCE must be 'balanced' or 'roundrobin'")
    load = [0] * world
    owner = [0] * n
    for i in sorted(range(n), key=lambda k: (-int(sizes[k]), k)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += int(sizes[i])
    return owner


def any_rank_failed(failed):
    """True on every rank if `failed` is true on any (one small all_reduce): a rank that hit an
    error must not leave the others waiting at the barrier until the collective times out."""
    import torch
    import torch.distributed as dist
    rank, world = init_process_group()
    if world <= 1:
        return bool(failed)
    dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    flag = torch.tensor([1 if failed else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    return bool(int(flag.item()))


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
