"""Multi-GPU driver of the hot path: one process per GPU (torchrun), `torch.distributed` for
the two collectives SURVEY.md §8(e) asks for.

    rank r holds 1/N of the reads
    1. k-mer records of the local reads, grouped by owner = (hash * N) >> 32    [device]
    2. ONE alltoallv routes every record to its owner                            [NCCL / NVLink]
    3. owner sorts, counts, filters -> its slice of the dictionary                [device]
    4. ONE allgather of the slices (rank order == global (hash, k-mer) order)     [NCCL / NVLink]
    5. prune / unipath edges / HBV replicated on every rank; ReadPaths of the local reads

The exchange helpers are backend agnostic (NCCL on GPUs, gloo on CPU tensors in the tests).
"""
from __future__ import annotations

import numpy as np
import torch

REC_WORDS = 4        # a k-mer record is 4 x u32
ENTRY_WORDS = 8      # a dictionary entry is 8 x u32


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


def owner_of(h, nparts):
    """Range partition of the 32-bit hash space: monotone in h."""
    return ((np.asarray(h, dtype=np.uint64) * np.uint64(nparts)) >> np.uint64(32)).astype(np.int64)


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
    n = dist.get_world_size()
    send_counts, send_ptr = ctx.mg_partition_records(params, n)
    recv_counts = exchange_counts(dist, send_counts, device)
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    recv_ptr = ctx.mg_recv_buffer(n_recv)
    send_t = dev_tensor(send_ptr, n_send * REC_WORDS, device)
    recv_t = dev_tensor(recv_ptr, n_recv * REC_WORDS, device)
    exchange_records(dist, send_t, send_counts, recv_t, recv_counts, REC_WORDS)
    torch.cuda.synchronize(device)
    n_k, dict_ptr = ctx.mg_count_received(n_recv)
    sizes_t = torch.zeros(n, dtype=torch.int64, device=device)
    sizes_t[dist.get_rank()] = n_k
    dist.all_reduce(sizes_t)
    sizes = [int(x) for x in sizes_t.tolist()]
    total = sum(sizes)
    full_ptr = ctx.mg_dictionary_buffer(total)
    full_t = dev_tensor(full_ptr, total * ENTRY_WORDS, device)
    local_t = dev_tensor(dict_ptr, n_k * ENTRY_WORDS, device)
    gather_slices(dist, local_t, full_t, sizes, ENTRY_WORDS)
    torch.cuda.synchronize(device)
    ctx.mg_install_dictionary(total)
    ctx.build_edges()
    ctx.build_hbv()
    if with_paths:
        ctx.path_reads()
    return dict(n_send=n_send, n_recv=n_recv, slice=n_k, total=total)
