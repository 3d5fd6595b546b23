"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path -- hash-range ownership,
the alltoallv of k-mer records and the uneven allgather of dictionary slices
(supernova_b200/multigpu.py) -- on records produced by the product's own extraction logic run
on the CPU (tests/hostsim).  The union must equal the oracle's dictionary of all reads."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reduce_numpy(recs):
    """count / ctx / barcode rule on records already owned by this rank -> entries sorted by (hash,kmer)."""
    from supernova_b200 import multigpu as mg
    if len(recs) == 0:
        return np.zeros((0, 8), np.uint32)
    h = mg.kmer_hash(recs[:, 0], recs[:, 1], recs[:, 2])
    order = np.lexsort((recs[:, 2], recs[:, 1], recs[:, 0], h))
    recs, h = recs[order], h[order]
    key = recs[:, :3]
    head = np.ones(len(recs), bool)
    head[1:] = (key[1:] != key[:-1]).any(axis=1)
    gid = np.cumsum(head) - 1
    cnt = np.bincount(gid)
    ctx = np.zeros(gid[-1] + 1, np.uint32)
    np.bitwise_or.at(ctx, gid, recs[:, 3] >> 24)
    bcv = (recs[:, 3] & 0xFFFFFF).astype(np.int64)
    mn = np.full(gid[-1] + 1, 1 << 30)
    mx = np.zeros(gid[-1] + 1, np.int64)
    pos = bcv > 0
    np.minimum.at(mn, gid[pos], bcv[pos])
    np.maximum.at(mx, gid[pos], bcv[pos])
    valid = (cnt >= 3) & (mx > 0) & (mn != mx)
    k = key[head][valid]
    out = np.zeros((len(k), 8), np.uint32)
    out[:, :3] = k
    out[:, 3] = np.minimum(cnt[valid], 0xFFFFFF) | (ctx[valid] << 24)
    out[:, 4] = 0xFFFFFFFF
    out[:, 6] = ctx[valid]
    out[:, 7] = h[head][valid]
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import datasets
    import supernova_b200 as sb
    from supernova_b200 import multigpu as mg
    from hostsim import lib
    from oracle.oracle import Oracle
    codes, quals, off, bc, _ = datasets.get("stress1")
    n = len(off) - 1
    lo, hi = n * rank // world, n * (rank + 1) // world            # this rank's reads
    o = Oracle(codes, quals, off, bc).stage("count")
    gl = o.good_len()
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=1)
    padded = np.concatenate([pb, np.zeros(32, np.uint8)])
    buf = np.zeros((256, 4), np.uint32)
    recs = [np.zeros((0, 4), np.uint32)]
    for r in range(lo, hi):
        m = lib().hs_extract_read(padded.ctypes.data + int(boff[r]), int(gl[r]), int(bc[r]), buf.ctypes.data)
        if m:
            recs.append(buf[:m].copy())
    recs = np.concatenate(recs)
    # 1. group by owner
    own = mg.owner_of(mg.kmer_hash(recs[:, 0], recs[:, 1], recs[:, 2]), world)
    order = np.argsort(own, kind="stable")
    recs, own = recs[order], own[order]
    send_counts = np.bincount(own, minlength=world).tolist()
    # 2. the alltoallv
    recv_counts = mg.exchange_counts(dist, send_counts, "cpu")
    send_t = torch.from_numpy(recs.astype(np.int32).ravel().copy())
    recv_t = torch.empty(sum(recv_counts) * mg.REC_WORDS, dtype=torch.int32)
    mg.exchange_records(dist, send_t, send_counts, recv_t, recv_counts, mg.REC_WORDS)
    got = recv_t.numpy().view(np.uint32).reshape(-1, 4)
    assert (mg.owner_of(mg.kmer_hash(got[:, 0], got[:, 1], got[:, 2]), world) == rank).all()
    # 3. this rank's slice
    sl = _reduce_numpy(got)
    # 4. the allgather
    sizes_t = torch.zeros(world, dtype=torch.int64)
    sizes_t[rank] = len(sl)
    dist.all_reduce(sizes_t)
    sizes = [int(x) for x in sizes_t.tolist()]
    full_t = torch.empty(sum(sizes) * mg.ENTRY_WORDS, dtype=torch.int32)
    mg.gather_slices(dist, torch.from_numpy(sl.astype(np.int32).ravel().copy()), full_t, sizes, mg.ENTRY_WORDS)
    full = full_t.numpy().view(np.uint32).reshape(-1, 8)
    # the gathered dictionary is globally ordered by (hash, k-mer) and equals the oracle's
    hk = full[:, 7].astype(np.uint64)
    assert (np.diff(hk.astype(np.int64)) >= 0).all()
    ok = o.kmers()
    idx = np.lexsort((full[:, 2], full[:, 1], full[:, 0]))
    res = np.array_equal(full[idx][:, :3], ok[:, :3]) and np.array_equal(full[idx][:, 3], ok[:, 3] | (ok[:, 4] << 24))
    q.put((rank, bool(res), len(full), sum(send_counts)))
    dist.destroy_process_group()


def test_two_rank_exchange_matches_oracle(built):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in out), out
    assert out[0][2] == out[1][2] > 0


def test_owner_is_monotone_and_balanced():
    sys.path.insert(0, ROOT)
    from supernova_b200 import multigpu as mg
    rng = np.random.default_rng(1)
    w = rng.integers(0, 2 ** 32, size=(200000, 3), dtype=np.uint64)
    h = mg.kmer_hash(w[:, 0], w[:, 1], w[:, 2])
    for n in (1, 2, 3, 8):
        o = mg.owner_of(h, n)
        assert o.min() >= 0 and o.max() <= n - 1
        srt = np.argsort(h)
        assert (np.diff(o[srt]) >= 0).all()
        cnt = np.bincount(o, minlength=n)
        assert cnt.max() < 1.05 * cnt.mean() + 50
