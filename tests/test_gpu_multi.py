"""-m gpu: the multi-GPU path (hash-range alltoallv + dictionary allgather over NCCL) must give
every rank the single-GPU dictionary / HBV and each rank the ReadPaths of its own reads.
World size = min(2, visible GPUs): with one GPU it still drives the partition / install path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, wd, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import datasets
        import supernova_b200 as sb
        from supernova_b200 import multigpu as mg
        codes, quals, off, bc, _ = datasets.get(name)
        n = len(off) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        a, b = int(off[lo]), int(off[hi])
        loff = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
        pb, boff, ln, pq, pqoff = sb.pack_reads(codes[a:b], quals[a:b], loff, threads=2)
        ctx = sb.Context(rank)
        ctx.load_reads(pb, boff, ln, pq, pqoff, bc[lo:hi])
        info = mg.build_distributed(ctx, dist, dev, sb.Params(), with_paths=True)
        km = ctx.kmers()
        ctx.write_hbv(os.path.join(wd, f"a{rank}.hbv"))
        po, poff, pe = ctx.paths()
        np.savez(os.path.join(wd, f"r{rank}.npz"), km=km, po=po, poff=poff, pe=pe, lo=lo, hi=hi)
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok", info))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error: " + repr(e) + traceback.format_exc(), None))


@pytest.mark.parametrize("name", ["stress2", "C1"])
def test_multi_gpu_matches_single(built, name, tmp_path):
    from oracle.oracle import Oracle
    import datasets
    world = min(2, torch.cuda.device_count())
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(s == "ok" for _, s, _ in out), out
    codes, quals, off, bc, _ = datasets.get(name)
    o = Oracle(codes, quals, off, bc).run()
    ok = o.kmers()
    o.write_hbv(str(tmp_path / "o.hbv"))
    oo, ooff, oe = o.paths()
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["km"][:, :3], ok[:, :3])
        assert np.array_equal(z["km"][:, 3], ok[:, 3] | (ok[:, 4] << 24))
        assert open(tmp_path / f"a{r}.hbv", "rb").read() == open(tmp_path / "o.hbv", "rb").read()
        lo, hi = int(z["lo"]), int(z["hi"])
        assert np.array_equal(z["po"], oo[lo:hi])
        assert np.array_equal(z["poff"] - z["poff"][0], ooff[lo:hi + 1] - ooff[lo])
        assert np.array_equal(z["pe"], oe[int(ooff[lo]):int(ooff[hi])])
