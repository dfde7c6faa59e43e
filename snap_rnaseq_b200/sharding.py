"""Host-side sharding and statistics reduction for multi-GPU runs.

The path shards by construction (reads are independent, the index is read-only and replicated per GPU), so there
is no collective on the data path: batches are dealt round-robin to ranks, each rank aligns its own batches, and the
only exchange is the end-of-run sum of the AlignerStats vector (SNAPLib/AlignerStats.cpp:75-102 does the same
per thread).  `torch.distributed` is used for that one all-reduce (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_ranges(n_items, world_size, rank, batch=1 << 17):
    """Round-robin deal of fixed-size batches: batch b goes to rank b % world_size.  Returns [(lo, hi), ...]."""
    out = []
    n_batches = (n_items + batch - 1) // batch
    for b in range(rank, n_batches, world_size):
        out.append((b * batch, min(n_items, (b + 1) * batch)))
    return out


def merge_sharded(results_by_rank, ranges_by_rank, n_items, dtype):
    """Inverse of shard_ranges: put each rank's results back in input order."""
    out = np.zeros(n_items, dtype)
    for res, ranges in zip(results_by_rank, ranges_by_rank):
        pos = 0
        for lo, hi in ranges:
            out[lo:hi] = res[pos:pos + hi - lo]
            pos += hi - lo
    return out


def allreduce_stats(stats_words, group=None):
    """Sum the flat int64 statistics vector (snapb200_stats) over all ranks.  No-op without torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(stats_words, dtype=np.int64).copy())
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def stats_from_results(paired_results):
    """The status/MAPQ part of the statistics vector computed on the host from result records (used by tests to
    check the device-side counters)."""
    from . import _abi as A
    w = np.zeros(A.STATS_WORDS, np.int64)
    st = paired_results["status"].ravel()
    mq = paired_results["mapq"].ravel()
    w[0] = st.size
    w[2] = int((st == A.SINGLE_HIT).sum())
    w[3] = int((st == A.MULTIPLE_HITS).sum())
    w[4] = int((st == A.NOT_FOUND).sum())
    w[6] = int(2 * paired_results["aligned_as_pair"].sum())
    ok = (st != A.NOT_FOUND) & (mq >= 0) & (mq <= 70)
    w[14:14 + 71] = np.bincount(mq[ok], minlength=71)[:71]
    return w


def fastq_shard_bounds(lib, text, world_size):
    """Byte ranges of one FASTQ text for `world_size` readers, as the reference splits a file over its worker threads: the file is
    cut at arbitrary offsets (RangeSplitter, SNAPLib/RangeSplitter.cpp:50-92) and a reader owns the records that START inside its
    range -- it skips to the first whole record (FASTQReader::skipPartialRecord = snapb200_fastq_record_start) and reads on past the
    end of its range to finish the last one.  Returns world_size + 1 offsets; rank r parses text[b[r]:b[r+1]], which holds exactly
    its records, with snapb200_fastq_parse.  Ranks whose cut lands after the last record start get an empty range."""
    text = memoryview(text)
    n = len(text)
    bounds = [0]
    for r in range(1, world_size):
        cut = n * r // world_size
        cut = max(cut, bounds[-1])
        # a record is assumed to fit the look-ahead window, as the reference assumes it fits its buffer (FASTQ.cpp:116-117)
        window = text[cut:cut + (1 << 20)]
        skip = lib.fastq_record_start(window) if cut < n else 0
        bounds.append(cut + skip if skip < len(window) or cut + len(window) < n else n)
    bounds.append(n)
    return bounds
