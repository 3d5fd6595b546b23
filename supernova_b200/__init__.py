"""supernova_b200 -- B200-native k-mer count -> unipath graph (HyperBasevector) -> ReadPath
hot path of 10x Genomics Supernova behind a C ABI (include/supernova_b200.h).

The compute lives in ``libsupernova_b200.so`` (hand-written sm_100a CUDA + C++ host code,
built by ``__graft_entry__.build()`` / ``make -C supernova_b200/csrc``); this package is the
thin ctypes mirror used by the tests and the bench.  There is no CPU fallback: creating a
``Context`` without a CUDA device raises.
"""
from .api import Context, Params, SnError, lib, pack_reads, write_read_files, pqvec_encode, pqvec_decode, nccl_unique_id, run_local_ranks  # noqa: F401
