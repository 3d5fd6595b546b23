"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path -- bucket-range ownership,
the alltoallv of super-k-mer records with their per-bucket counts and the uneven allgather of the
surviving k-mers (supernova_b200/multigpu.py) -- on records produced by the product's own MSP
logic run on the CPU (tests/hostsim).  The union must equal the oracle's dictionary of all reads."""
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
    out = np.zeros((len(k), 4), np.uint32)
    out[:, :3] = k
    out[:, 3] = np.minimum(cnt[valid], 0xFFFFFF) | (ctx[valid] << 24)
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
    padded = np.concatenate([pb, np.zeros(64, np.uint8)])
    buf = np.zeros((256, 4), np.uint32); bh = np.zeros(256, np.uint32); skb = np.zeros((256, 8), np.uint32); nsk = np.zeros(1, np.uint32)
    sk = [np.zeros((0, 8), np.uint32)]
    for r in range(lo, hi):
        m = lib().hs_msp_read(padded.ctypes.data + int(boff[r]), int(gl[r]), int(bc[r]), buf.ctypes.data, bh.ctypes.data, skb.ctypes.data, nsk.ctypes.data)
        if m:
            sk.append(skb[:int(nsk[0])].copy())
    sk = np.concatenate(sk)
    # 1. bucket order; owner = bucket * world >> bits
    bits = mg.bucket_bits(o.n_occ)
    bkt = (sk[:, 1] >> np.uint32(32 - bits)).astype(np.int64)
    order = np.argsort(bkt, kind="stable")
    sk, bkt = sk[order], bkt[order]
    hist = np.bincount(bkt, minlength=1 << bits).astype(np.int32)
    fb = [mg.first_bucket(w, world, bits) for w in range(world + 1)]
    assert (mg.owner_of_bucket(bkt, world, bits) == np.searchsorted(np.array(fb[1:]), bkt, side="right")).all()
    send_counts = [int(hist[fb[w]:fb[w + 1]].sum()) for w in range(world)]
    nbl = fb[rank + 1] - fb[rank]
    # 2. the alltoallv: records, and the per-bucket counts of the same bucket ranges
    recv_counts = mg.exchange_counts(dist, send_counts, "cpu")
    send_t = torch.from_numpy(sk.astype(np.int32).ravel().copy())
    recv_t = torch.empty(sum(recv_counts) * mg.SK_WORDS, dtype=torch.int32)
    mg.exchange_records(dist, send_t, send_counts, recv_t, recv_counts, mg.SK_WORDS)
    cnt_recv = torch.empty(world * nbl, dtype=torch.int32)
    dist.all_to_all_single(cnt_recv, torch.from_numpy(hist.copy()), output_split_sizes=[nbl] * world,
                           input_split_sizes=[fb[w + 1] - fb[w] for w in range(world)])
    got = recv_t.numpy().view(np.uint32).reshape(-1, 8)
    gb = (got[:, 1] >> np.uint32(32 - bits)).astype(np.int64)
    assert ((gb >= fb[rank]) & (gb < fb[rank + 1])).all()
    # the received counts, scanned segment-major, delimit every (source, bucket) range of the receive buffer
    off = np.concatenate([[0], np.cumsum(cnt_recv.numpy().astype(np.int64))])
    assert off[-1] == len(got)
    for s_ in range(world):
        for j in (0, nbl // 2, nbl - 1):
            seg = gb[off[s_ * nbl + j]:off[s_ * nbl + j + 1]]
            assert (seg == fb[rank] + j).all()
    # 3. this rank's surviving k-mers
    krec = np.zeros((int((((got[:, 0] >> 24) & 0x3F) + 1).sum()), 4), np.uint32)
    m = lib().hs_sk_expand(np.ascontiguousarray(got).ctypes.data, len(got), krec.ctypes.data)
    assert m == len(krec)
    sl = _reduce_numpy(krec)
    # 4. the allgather
    sizes_t = torch.zeros(world, dtype=torch.int64)
    sizes_t[rank] = len(sl)
    dist.all_reduce(sizes_t)
    sizes = [int(x) for x in sizes_t.tolist()]
    full_t = torch.empty(sum(sizes) * mg.SURV_WORDS, dtype=torch.int32)
    mg.gather_slices(dist, torch.from_numpy(sl.astype(np.int32).ravel().copy()), full_t, sizes, mg.SURV_WORDS)
    full = full_t.numpy().view(np.uint32).reshape(-1, 4)
    # the gathered k-mers are the oracle's dictionary
    ok = o.kmers()
    idx = np.lexsort((full[:, 2], full[:, 1], full[:, 0]))
    res = np.array_equal(full[idx][:, :3], ok[:, :3]) and np.array_equal(full[idx][:, 3], ok[:, 3] | (ok[:, 4] << 24))
    q.put((rank, bool(res), len(full), sum(send_counts)))
    dist.destroy_process_group()


def _worker_passes(rank, world, port, q, lp):
    """The sharded count in 2^lp interleaved passes (sn_multi.cu: mg_count_sharded), host arithmetic only: per pass the records of
    msp_window_bucket's window, renumbered owner-major, exchanged and reduced; a rank's per-pass results in pass order must be in
    global bucket order, and the union over the ranks the oracle's dictionary."""
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
    lo, hi = n * rank // world, n * (rank + 1) // world
    o = Oracle(codes, quals, off, bc).stage("count")
    gl = o.good_len()
    pb, boff, pl, pq, pqoff = sb.pack_reads(codes, quals, off, threads=1)
    padded = np.concatenate([pb, np.zeros(64, np.uint8)])
    buf = np.zeros((256, 4), np.uint32); bh = np.zeros(256, np.uint32); skb = np.zeros((256, 8), np.uint32); nsk = np.zeros(1, np.uint32)
    sk = [np.zeros((0, 8), np.uint32)]
    for r in range(lo, hi):
        m = lib().hs_msp_read(padded.ctypes.data + int(boff[r]), int(gl[r]), int(bc[r]), buf.ctypes.data, bh.ctypes.data, skb.ctypes.data, nsk.ctypes.data)
        if m:
            sk.append(skb[:int(nsk[0])].copy())
    sk = np.concatenate(sk)
    lnr = world.bit_length() - 1
    bits = max(mg.bucket_bits(o.n_occ), lnr + lp + 1)
    wb = bits - lnr
    gbkt = (sk[:, 1] >> np.uint32(32 - bits)).astype(np.int64)
    per_pass = (1 << bits) >> lp
    nbl = 1 << (wb - lp)                                              # buckets a rank receives per pass
    mine, mine_bucket = [], []
    for ps in range(1 << lp):
        pcfg = wb | (lp << 8) | (ps << 16)
        table = np.array([lib().hs_window_bucket(b, 0, per_pass, pcfg) for b in range(1 << bits)], np.int64)
        comp = table[gbkt]
        sel = comp != 0xFFFFFFFF
        s_p, c_p, g_p = sk[sel], comp[sel], gbkt[sel]
        order = np.argsort(c_p, kind="stable")
        s_p, c_p, g_p = s_p[order], c_p[order], g_p[order]
        send_counts = [int(((c_p >> (wb - lp)) == w).sum()) for w in range(world)]
        recv_counts = mg.exchange_counts(dist, send_counts, "cpu")
        send_t = torch.from_numpy(s_p.astype(np.int32).ravel().copy())
        recv_t = torch.empty(sum(recv_counts) * mg.SK_WORDS, dtype=torch.int32)
        mg.exchange_records(dist, send_t, send_counts, recv_t, recv_counts, mg.SK_WORDS)
        got = recv_t.numpy().view(np.uint32).reshape(-1, 8)
        gb = (got[:, 1] >> np.uint32(32 - bits)).astype(np.int64)
        # everything received belongs to this rank (owner = top bits) and to this pass (next lp bits)
        assert ((gb >> wb) == rank).all() and (((gb & ((1 << wb) - 1)) >> (wb - lp)) == ps).all()
        krec = np.zeros((int((((got[:, 0] >> 24) & 0x3F) + 1).sum()), 4), np.uint32)
        m = lib().hs_sk_expand(np.ascontiguousarray(got).ctypes.data, len(got), krec.ctypes.data) if len(got) else 0
        assert m == len(krec)
        mine.append(_reduce_numpy(krec))
        mine_bucket.append((rank << wb) + (ps << (wb - lp)))          # first global bucket of this slice
    assert mine_bucket == sorted(mine_bucket) and mine_bucket[0] == rank << wb      # pass order is bucket order inside the rank's window
    sl = np.concatenate(mine)
    sizes_t = torch.zeros(world, dtype=torch.int64)
    sizes_t[rank] = len(sl)
    dist.all_reduce(sizes_t)
    sizes = [int(x) for x in sizes_t.tolist()]
    full_t = torch.empty(sum(sizes) * mg.SURV_WORDS, dtype=torch.int32)
    mg.gather_slices(dist, torch.from_numpy(sl.astype(np.int32).ravel().copy()), full_t, sizes, mg.SURV_WORDS)
    full = full_t.numpy().view(np.uint32).reshape(-1, 4)
    ok = o.kmers()
    idx = np.lexsort((full[:, 2], full[:, 1], full[:, 0]))
    res = len(full) == len(ok) and np.array_equal(full[idx][:, :3], ok[:, :3]) and np.array_equal(full[idx][:, 3], ok[:, 3] | (ok[:, 4] << 24))
    q.put((rank, bool(res), len(full), nbl))
    dist.destroy_process_group()


@pytest.mark.parametrize("lp", [1, 2])
def test_two_rank_count_in_interleaved_passes_matches_oracle(built, lp):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_passes, args=(r, world, port, q, lp)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in out), out
    assert out[0][2] == out[1][2] > 0


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


def test_bucket_ownership_is_a_balanced_range_partition(built):
    sys.path.insert(0, ROOT)
    import supernova_b200 as sb
    from supernova_b200 import multigpu as mg
    for n_occ in (0, 1, 1536 << 4, (1536 << 4) + 16, 790_041_546, 1 << 40, 1 << 60):
        assert sb.lib().sn_msp_bucket_bits(n_occ) == mg.bucket_bits(n_occ)
    for bits in (4, 10, 19):
        b = np.arange(1 << bits)
        for n in (1, 2, 3, 8):
            o = mg.owner_of_bucket(b, n, bits)
            assert o.min() == 0 and o.max() == n - 1 and (np.diff(o.astype(np.int64)) >= 0).all()
            fb = [mg.first_bucket(w, n, bits) for w in range(n + 1)]
            assert fb[0] == 0 and fb[-1] == 1 << bits
            for w in range(n):
                assert (o[fb[w]:fb[w + 1]] == w).all()
            cnt = np.bincount(o.astype(np.int64), minlength=n)
            assert cnt.max() - cnt.min() <= 1
    assert mg.bucket_bits(0) == 4 and mg.bucket_bits(790_041_546) == 19 and mg.bucket_bits(1 << 60) == 24
