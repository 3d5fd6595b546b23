"""Multi-GPU driver of the hot path: one process per GPU (torchrun), `torch.distributed` for
the two collectives SURVEY.md §8(e) asks for.

    rank r holds 1/N of the reads
    1. good lengths; the k-mer occurrences of all ranks fix the number of minimizer buckets  [allreduce of one number]
    2. super-k-mer records of the local reads in bucket order (MSP)                          [device]
    3. ONE alltoallv routes every record to the owner of its bucket,
       owner(bucket) = bucket * N >> bits (plus the per-bucket counts of the same ranges)    [NCCL / NVLink]
    4. owner counts and filters its buckets -> its surviving k-mers, bucket order             [device]
    5. ONE allgather of the survivors (+ per-bucket counts): rank order is bucket order, so the
       gathered k-mers ARE the dictionary                                                    [NCCL / NVLink]
    6. prune / unipath edges / HBV replicated on every rank; ReadPaths of the local reads

A super-k-mer record is 32 bytes for ~14 k-mers, so the exchange moves ~2.3 bytes per k-mer
occurrence (a k-mer record would be 16).  The exchange helpers are backend agnostic (NCCL on
GPUs, gloo on CPU tensors in the tests).
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


def build_distributed(ctx, dist, device, params=None, with_paths=False):
    """Runs the hot path for this rank's context; returns the per-rank counts dict."""
    from .api import Params
    params = params or Params()
    n, rank = dist.get_world_size(), dist.get_rank()
    # 1. good lengths; global number of k-mer occurrences -> bucket bits (equal on every rank)
    occ = torch.tensor([ctx.mg_good_lengths(params)], dtype=torch.int64, device=device)
    dist.all_reduce(occ)
    bits = bucket_bits(int(occ.item()))
    while (1 << bits) < n:
        bits += 1
    # 2. super-k-mers of the local reads, bucket order
    send_counts, rec_ptr, cnt_ptr = ctx.mg_partition(bits, n)
    fb = [first_bucket(o, n, bits) for o in range(n + 1)]
    nbl = fb[rank + 1] - fb[rank]
    # 3. the alltoallv (records + the per-bucket counts of the same bucket ranges)
    recv_counts = exchange_counts(dist, send_counts, device)
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    send_t = dev_tensor(rec_ptr, n_send * SK_WORDS, device)
    recv_t = dev_tensor(ctx.mg_recv_records(n_recv), n_recv * SK_WORDS, device)
    exchange_records(dist, send_t, send_counts, recv_t, recv_counts, SK_WORDS)
    cnt_send = dev_tensor(cnt_ptr, 1 << bits, device)
    cnt_recv = dev_tensor(ctx.mg_recv_counts(n * nbl), n * nbl, device)
    dist.all_to_all_single(cnt_recv, cnt_send, output_split_sizes=[nbl] * n, input_split_sizes=[fb[o + 1] - fb[o] for o in range(n)])
    torch.cuda.synchronize(device)
    # 4. count + filter of this rank's buckets
    n_s, surv_ptr, scnt_ptr = ctx.mg_count_received(n, nbl, n_recv)
    # 5. the allgather of the surviving k-mers
    sizes_t = torch.zeros(n, dtype=torch.int64, device=device)
    sizes_t[rank] = n_s
    dist.all_reduce(sizes_t)
    sizes = [int(x) for x in sizes_t.tolist()]
    total = sum(sizes)
    full_t = dev_tensor(ctx.mg_survivor_buffer(total), total * SURV_WORDS, device)
    local_t = dev_tensor(surv_ptr, n_s * SURV_WORDS, device)
    gather_slices(dist, local_t, full_t, sizes, SURV_WORDS)
    gcnt_t = dev_tensor(ctx.mg_bucket_count_buffer(bits), 1 << bits, device)
    gather_slices(dist, dev_tensor(scnt_ptr, nbl, device), gcnt_t, [fb[o + 1] - fb[o] for o in range(n)], 1)
    torch.cuda.synchronize(device)
    ctx.mg_install_survivors(total, bits)
    # 6. the graph, replicated
    ctx.build_edges()
    ctx.build_hbv()
    if with_paths:
        ctx.path_reads()
    return dict(bits=bits, n_send=n_send, n_recv=n_recv, survivors=n_s, total=total)
