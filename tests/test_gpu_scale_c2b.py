"""-m gpu: C2b -- BASELINE.json's "56x of a 63 Mbp diploid" taken literally (3.53 Gbp, 2.3e9 k-mer occurrences, an HBV of
two giant strand components) -- EXACT parity against the reference's own buildReadQGraph48 run through the committed
golden digests (tests/golden/scale_digests.json; see tests/test_gpu_scale.py).  In its own module so that its 30 GB of
host arrays never coexist with the C2 fixture.  Bar: bit-exact."""
import os

import numpy as np
import pytest

from test_gpu_scale import GOLD, against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(built):
    import supernova_b200
    return supernova_b200


def test_c2b_matches_the_reference_exactly(sb, tmp_path):
    """C2b: "56x of a 63 Mbp diploid" taken literally -- 3.53 Gbp, 2.3e9 k-mer occurrences, two giant strand
    components in the HBV -- against the reference's own run (golden digests)."""
    from supernova_b200 import synth
    g = GOLD["C2b"]
    b, q, bc, _ = synth.make_reads(g["G"], g["pairs"], g["n_bc"], g["seed"], workers=min(os.cpu_count() or 1, 32))
    n, L = b.shape
    off = np.arange(n + 1, dtype=np.uint64) * L
    packed = sb.pack_reads(b.ravel(), q.ravel(), off)
    del b, q
    wd = str(tmp_path)
    with sb.Context(0) as ctx:
        ctx.load_reads(*packed, bc)
        ctx.build_read_qgraph48(wd, sb.Params(), with_paths=True)
        against_golden(sb, ctx, "C2b", wd, packed, bc)


