"""Host-side logic that needs no GPU: the read simulator, batch containers, round-robin sharding and the
statistics all-reduce (gloo, world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from snap_rnaseq_b200 import _abi as A  # noqa: E402
from snap_rnaseq_b200 import sharding, synth  # noqa: E402


def test_batch_roundtrip_and_slices():
    seqs = ["ACGT", "", "NNNN", "A" * 37]
    b = A.Batch.from_strings(seqs)
    assert b.n == 4 and [b.read(i)[0] for i in range(4)] == seqs
    s = b.slice(1, 4)
    assert s.n == 3 and s.read(2)[0] == "A" * 37 and int(s.offsets[0]) == 0
    e = A.Batch.from_strings([])
    assert e.n == 0 and e.offsets.tolist() == [0]


def test_simulator_is_seeded_and_plausible():
    contigs = synth.random_contigs([30000, 20000], seed=20)
    a = synth.simulate(contigs, 500, 100, paired=True, err=0.02, seed=5)
    b = synth.simulate(contigs, 500, 100, paired=True, err=0.02, seed=5)
    assert np.array_equal(a["batches"][0].bases, b["batches"][0].bases) and np.array_equal(a["batches"][1].quals, b["batches"][1].quals)
    c = synth.simulate(contigs, 500, 100, paired=True, err=0.02, seed=6)
    assert not np.array_equal(a["batches"][0].bases, c["batches"][0].bases)
    # error-free forward-strand reads are substrings of their contig; mates are reverse complements of the far end
    z = synth.simulate(contigs, 200, 80, paired=True, err=0.0, indel_frac=0.0, n_rate=0.0, seed=9)
    names = z["names"]
    for i in range(200):
        r1, _ = z["batches"][0].read(i)
        r2, _ = z["batches"][1].read(i)
        ref = contigs[names[z["contig"][i]]].tobytes().decode()
        s0, s1 = int(z["start"][i][0]), int(z["start"][i][1])
        left, right = ref[s0:s0 + 80], ref[s1:s1 + 80]
        rc = right[::-1].translate(str.maketrans("ACGT", "TGCA"))
        assert (r1, r2) == ((left, rc) if z["strand"][i] == 0 else (rc, left))
    # qualities: >= 90 % of bases at >= Q20 (the reference's default quality gate, Read.h:422-433)
    q = a["batches"][0].quals.astype(int) - 33
    assert (q >= 20).mean() > 0.9


def test_snap_layout_matches_reference_padding():
    contigs = synth.random_contigs([100, 50], seed=1)
    bases, offs = synth.snap_layout(contigs, 500)
    assert offs.tolist() == [500, 1100] and bases.size == 500 + 100 + 500 + 50 + 500
    assert bytes(bases[:500]) == b"n" * 500 and bytes(bases[-500:]) == b"n" * 500
    assert bytes(bases[500:600]) == contigs["chr1"].tobytes()


def test_round_robin_sharding_covers_everything_once():
    n = 1_000_003
    for world in (1, 2, 3, 8):
        seen = np.zeros(n, np.int32)
        for r in range(world):
            for lo, hi in sharding.shard_ranges(n, world, r, batch=1 << 16):
                seen[lo:hi] += 1
        assert (seen == 1).all()
    # merge restores input order
    data = np.arange(1000, dtype=np.int64)
    ranges = [sharding.shard_ranges(1000, 3, r, batch=64) for r in range(3)]
    parts = [np.concatenate([data[lo:hi] for lo, hi in rg]) for rg in ranges]
    assert np.array_equal(sharding.merge_sharded(parts, ranges, 1000, np.int64), data)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank "aligns" its shard: here the results are synthetic records derived from the item index
    n = 10_000
    res = np.zeros(n, A.PAIRED_RESULT)
    idx = np.arange(n)
    res["status"][:, 0] = idx % 3
    res["status"][:, 1] = (idx // 3) % 3
    res["mapq"][:, 0] = idx % 71
    res["mapq"][:, 1] = (idx * 7) % 71
    res["aligned_as_pair"] = (idx % 5 != 0)
    mine = np.concatenate([res[lo:hi] for lo, hi in sharding.shard_ranges(n, world, rank, batch=512)])
    total = sharding.allreduce_stats(sharding.stats_from_results(mine))
    q.put((rank, total, sharding.stats_from_results(res)))
    dist.destroy_process_group()


def test_stats_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, expect in got:
        assert np.array_equal(total, expect), f"rank {rank}: reduced stats differ from the unsharded ones"


def _fastq_worker(rank, world, port, q):
    """Each rank takes its byte range of one FASTQ text (found with the product's snapb200_fastq_record_start) and parses it; the
    per-record logic runs through tests/hostsim here because this box has no GPU (on a GPU box the same ranges go to
    snapb200_fastq_parse).  The only exchange is the read count (gloo all-reduce), as for the statistics."""
    import ctypes
    import subprocess
    import torch
    import torch.distributed as dist
    import snap_rnaseq_b200 as S
    from snap_rnaseq_b200._binding import BatchLib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import io_cases
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    if rank == 0 and not os.path.exists(so):
        subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    dist.barrier()
    hostsim = BatchLib(ctypes.CDLL(so), "hostsim_")
    text = io_cases.fastq_text(5, 400, 100, crlf_frac=0.1, partial_tail=False)
    bounds = sharding.fastq_shard_bounds(S.lib(), text, world)
    mine, used = hostsim.fastq_parse(text[bounds[rank]:bounds[rank + 1]], 3)
    assert used == bounds[rank + 1] - bounds[rank]
    t = torch.tensor([mine.n], dtype=torch.int64)
    dist.all_reduce(t)
    q.put((rank, int(t[0]), bounds, mine.ids[:mine.id_offsets[-1]].tobytes(), mine.clipped_len[:mine.n].tolist()))
    dist.destroy_process_group()


def test_fastq_sharding_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fastq_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] == 400 and got[0][2] == got[1][2]
    bounds = got[0][2]
    assert 0 < bounds[1] < bounds[2] and len(got[0][4]) > 100 and len(got[1][4]) > 100
    # the shards, in rank order, are the unsharded parse
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ctypes
    import io_cases
    from snap_rnaseq_b200._binding import BatchLib
    hostsim = BatchLib(ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim", "libiohostsim.so")), "hostsim_")
    whole, _ = hostsim.fastq_parse(io_cases.fastq_text(5, 400, 100, crlf_frac=0.1, partial_tail=False), 3)
    assert got[0][3] + got[1][3] == whole.ids[:whole.id_offsets[-1]].tobytes()
    assert got[0][4] + got[1][4] == whole.clipped_len[:whole.n].tolist()
