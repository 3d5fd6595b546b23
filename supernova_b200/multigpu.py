"""Multi-GPU launcher glue: one process per GPU (torchrun).  The data path is csrc/sn_multi.cu -- the C++ host issues
every collective on NCCL (sn_comm.cu); `torch.distributed` is used here only to hand the NCCL unique id to the ranks
(and by bench.py for its barrier and the max-over-ranks of the timings).

    rank r holds 1/N of the reads
    1. good lengths; the k-mer occurrences of all ranks fix the number of minimizer buckets
    2. super-k-mer records of the local reads in bucket order (MSP)
    3. ONE alltoallv routes every record to the owner of its bucket, owner(bucket) = bucket * N >> bits
    4. owner counts and filters its buckets -> its shard of the dictionary (it STAYS sharded)
    5. neighbours on other ranks: ghost query/answer exchange; recomputeAdjacencies; links; chains cut at rank borders
    6. the stop table of all ranks is gathered, every rank derives all edge lengths / ids / offsets; bases by all-reduce
    7. HyperBasevector on every rank; ReadPaths of the local reads (the finished k-mer table is gathered for them)

The helpers below (bucket ownership, exchange with uneven sizes) are the numpy/torch twins of the C++ logic, used by the
world-size-2 gloo tests on CPU (tests/test_multi_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch

SK_WORDS = 8         # a super-k-mer record is 8 x u32
SURV_WORDS = 4       # a surviving k-mer is 4 x u32: w0, w1, w2, count:24 | ctx << 24


def kmer_hash(w0, w1, w2):
    """numpy twin of sn::kmer_hash (supernova_b200/csrc/sn_kmer.cuh) on uint64 arrays."""
    M = np.uint64(0xFFFFFFFF)
    w0, w1, w2 = (np.asarray(x, dtype=np.uint64) for x in (w0, w1, w2))
    h = (w0 * np.uint64(0x9E3779B1)) & M
    h = (((h ^ (h >> np.uint64(15))) + w1) * np.uint64(0x85EBCA77)) & M
    h = (((h ^ (h >> np.uint64(13))) + w2) * np.uint64(0xC2B2AE3D)) & M
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & M
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & M
    h ^= h >> np.uint64(16)
    return h


def bucket_bits(n_occ_total):
    """python twin of sn::msp_bucket_bits: 768..1536 k-mer occurrences per bucket."""
    b = 4
    while b < 24 and (int(n_occ_total) >> b) > 1536:
        b += 1
    return b


def first_bucket(owner, nparts, bits):
    """owner(bucket) = bucket * nparts >> bits  <=>  owner o holds [first_bucket(o), first_bucket(o+1))."""
    return ((owner << bits) + nparts - 1) // nparts


def owner_of_bucket(bucket, nparts, bits):
    return (np.asarray(bucket, dtype=np.uint64) * np.uint64(nparts)) >> np.uint64(bits)


class _DevArray:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def dev_tensor(ptr, n_words, device):
    if n_words == 0:
        return torch.empty(0, dtype=torch.int32, device=device)
    return torch.as_tensor(_DevArray(ptr, n_words), device=device)


def exchange_counts(dist, send_counts, device):
    """alltoall of the per-destination record counts -> per-source counts."""
    s = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    return [int(x) for x in r.tolist()]


def exchange_records(dist, send, send_counts, recv, recv_counts, words):
    """THE alltoallv: `send` holds the records grouped by destination rank."""
    dist.all_to_all_single(recv, send, output_split_sizes=[c * words for c in recv_counts],
                           input_split_sizes=[c * words for c in send_counts])


def gather_slices(dist, local, full, sizes, words):
    """THE allgather with uneven slices: rank r's slice goes to full[off_r : off_r + n_r]."""
    off = 0
    rank = dist.get_rank()
    for r, n in enumerate(sizes):
        sl = full[off * words:(off + n) * words]
        if r == rank and n:
            sl.copy_(local[:n * words])
        if n:
            dist.broadcast(sl, src=r)
        off += n


def init_comm(ctx, dist, device):
    """One NCCL communicator per context, created by the library itself (csrc/sn_comm.cu: ncclCommInitRank): rank 0
    makes the unique id, `torch.distributed` only carries those 128 bytes to the other ranks."""
    if getattr(ctx, "_comm_world", None) == dist.get_world_size():
        return
    from .api import nccl_unique_id
    n, rank = dist.get_world_size(), dist.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(uid, src=0)
    ctx.comm_init_nccl(rank, n, bytes(uid.cpu().numpy().tobytes()))
    ctx._comm_world = n


def build_distributed(ctx, dist, device, params=None, with_paths=False):
    """The hot path for this rank's context over all ranks: sn_mg_build_graph (csrc/sn_multi.cu) -- every collective
    (the alltoallv of super-k-mer records, the ghost query/answer exchanges, the gathers of the stop table, the
    all-reduce of the edge bases) is issued by the C++ host on NCCL; nothing of the data path runs through Python."""
    from .api import Params
    init_comm(ctx, dist, device)
    ctx.mg_build_graph(params or Params(), with_paths=with_paths)
    c = ctx.counts()
    return dict(sharded=ctx.dict_is_sharded(), n_kmers_local=c["n_kmers"], n_edges=c["n_edges"])
