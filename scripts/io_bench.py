#!/usr/bin/env python3
"""Throughput of the two I/O-edge stages on one GPU, with the reference's own code timed beside them.

  FASTQ text of both mates --snapb200_fastq_parse--> read arrays --snapb200_paired_batch--> alignments
                           --snapb200_sam_batch--> SAM text

Prints one JSON line: per stage the kernel time (CUDA events, no copies), the end-to-end time of the C-ABI call from host
buffers, the algorithmic bytes (text in + arrays out; arrays + genome windows in + text out) against the measured HBM copy
peak, and the compiled reference (FASTQReader::getNextRead, SimpleReadWriter::writePair; one thread, as each of the
reference's worker threads runs them) on a sample of the same bytes -- which is also the parity check: the sample's arrays
and SAM text must be identical.  TEST INFRASTRUCTURE touches oracle/ only for that check.
usage: io_bench.py [pairs] [genome_mbp] [read_len]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import snap_rnaseq_b200 as S  # noqa: E402
from snap_rnaseq_b200 import _abi as A, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mbp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rlen = int(sys.argv[3]) if len(sys.argv) > 3 else 150
SAMPLE = 1_000_000  # pairs the compiled reference re-does for the parity check and its own timing (one thread: a few seconds)


def main():
    L = S.lib(0)
    bench.GENOME_CONTIGS, bench.READ_LEN, bench.ERR_RATE = [25_000_000] * max(1, mbp // 25), rlen, 0.01
    contigs = bench.make_genome()
    bases, offs = synth.snap_layout(contigs, 500)
    h = L.build_index(bases, offs, piece_names=list(contigs), seed_len=20)
    b0, b1 = bench.make_pairs(contigs, pairs, seed=77)
    texts = [synth.fastq_fixed(b0, 0), synth.fastq_fixed(b1, 1)]
    peak, peak_src = bench.measured_peaks()
    out = {"pairs": pairs, "genome_mbp": mbp, "read_len": rlen, "hbm_peak_gbs": peak, "hbm_peak_source": peak_src}

    # ---- FASTQ parse ---------------------------------------------------------------------------------------------------
    import torch

    def pinned(n, dtype):
        return torch.empty(n, dtype=dtype).pin_memory().numpy()

    texts = [np.copyto(p := pinned(t.size, torch.uint8), t) or p for t in texts]
    nb = texts[0].size
    bufs = [(pinned(pairs + 1, torch.int32).view(np.uint32), pinned(pairs + 1, torch.int32).view(np.uint32), pinned(nb, torch.uint8),
             pinned(nb, torch.uint8), pinned(nb, torch.uint8), pinned(pairs, torch.int16).view(np.uint16), pinned(pairs, torch.int16).view(np.uint16))
            for _ in range(2)]
    reads = []
    k_ms = e2e_s = 0.0
    for rep in range(3):  # the last repetition counts (buffers allocated, clocks up)
        reads, k_ms, e2e_s = [], 0.0, 0.0
        for t, bf in zip(texts, bufs):
            t0 = time.perf_counter()
            r, used = L.fastq_parse(t, 3, bufs=bf)
            e2e_s += time.perf_counter() - t0
            k_ms += L.io_last_kernel_ms()[0]
            assert used == t.size and r.n == pairs
            reads.append(r)
    text_bytes = sum(t.size for t in texts)
    arrays = sum(int(r.offsets[-1]) * 2 + int(r.id_offsets[-1]) + r.n * (4 + 4 + 2 + 2) for r in reads)
    algo = text_bytes + arrays  # every text byte read once, every array byte written once
    out["fastq_parse"] = {"reads": 2 * pairs, "text_bytes": text_bytes, "kernel_ms": k_ms, "e2e_ms": e2e_s * 1e3,
                          "reads_per_s_kernels": 2 * pairs / (k_ms * 1e-3), "reads_per_s_e2e": 2 * pairs / e2e_s,
                          "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / (k_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": algo,
                                       "kernels": "fq_count_kernel, fq_positions_kernel, fq_record_kernel, fq_copy_kernel + 3 cub scans"}}

    # ---- align (the hot path; its numbers of record are bench.py's) -------------------------------------------------------
    pp = A.paired_defaults()
    c0, c1 = reads[0].clipped_batch(), reads[1].clipped_batch()
    t0 = time.perf_counter()
    res = L.paired(h, pp, c0, c1)
    out["align_e2e_ms"] = (time.perf_counter() - t0) * 1e3
    aln = []
    for e in range(2):
        a = np.zeros(pairs, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction"):
            a[f] = res[f][:, e]
        aln.append(a)

    # ---- SAM text -------------------------------------------------------------------------------------------------------
    buf = pinned(2 * pairs * (2 * rlen + 220), torch.uint8)
    for rep in range(3):
        t0 = time.perf_counter()
        sam, lo = L.sam(h, reads[0], reads[1], aln[0], aln[1], False, None, out=buf)
        e2e_s = time.perf_counter() - t0
        k_ms = L.io_last_kernel_ms()[1]
    mapped = int((aln[0]["status"] != 0).sum() + (aln[1]["status"] != 0).sum())
    in_bytes = sum(int(r.offsets[-1]) * 2 + int(r.id_offsets[-1]) + r.n * (4 + 4 + 2 + 2 + 12) for r in reads) + mapped * (rlen + 2 * 40)
    algo = in_bytes + int(lo[-1])
    out["sam_text"] = {"lines": 2 * pairs, "sam_bytes": int(lo[-1]), "kernel_ms": k_ms, "e2e_ms": e2e_s * 1e3,
                       "reads_per_s_kernels": 2 * pairs / (k_ms * 1e-3), "reads_per_s_e2e": 2 * pairs / e2e_s,
                       "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": algo / (k_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": algo,
                                    "kernels": "sam_measure_kernel (CIGAR by Landau-Vishkin + line lengths), cub scan, sam_write_kernel"}}

    # ---- BAM records (the same kernels, SNAPB200_SAM_BAM_RECORDS) -----------------------------------------------------------
    for rep in range(3):
        t0 = time.perf_counter()
        bam, blo = L.sam(h, reads[0], reads[1], aln[0], aln[1], False, None, out=buf, bam=True)
        e2e_b = time.perf_counter() - t0
        k_b = L.io_last_kernel_ms()[1]
    out["bam_records"] = {"records": 2 * pairs, "bam_bytes": int(blo[-1]), "kernel_ms": k_b, "e2e_ms": e2e_b * 1e3,
                          "reads_per_s_kernels": 2 * pairs / (k_b * 1e-3), "reads_per_s_e2e": 2 * pairs / e2e_b,
                          "note": "uncompressed BAMFormat::writeRead records; BGZF is the reference's host filter"}

    # ---- BGZF container over those BAM records (row f4b; the reference: zlib level 6 per 64 KiB chunk on its host threads) ----------
    import zlib
    bam_bytes = bytes(bam[:int(blo[-1])]) if not isinstance(bam, bytes) else bam
    for rep in range(2):
        t0 = time.perf_counter()
        z, zoff = L.bgzf_compress(bam_bytes)
        e2e_z = time.perf_counter() - t0
        k_z = L.bgzf_last_kernel_ms()
    sample = bam_bytes[:64 << 20]
    t0 = time.perf_counter()
    zs = sum(len(zlib.compress(sample[i:i + 65536], 6)) for i in range(0, len(sample), 65536))
    t_zlib = time.perf_counter() - t0
    import gzip
    out["bgzf"] = {"input_bytes": len(bam_bytes), "output_bytes": len(z), "ratio": len(z) / len(bam_bytes), "blocks": len(zoff) - 1, "kernel_ms": k_z,
                   "e2e_ms": e2e_z * 1e3, "gb_per_s_kernels": len(bam_bytes) / (k_z * 1e-3) / 1e9, "gb_per_s_e2e": len(bam_bytes) / e2e_z / 1e9,
                   "inflates_to_input": gzip.decompress(z[:int(zoff[200])]) == bam_bytes[:200 * 65024],
                   "zlib_level6_one_thread": {"sample_bytes": len(sample), "ratio": zs / len(sample), "gb_per_s": len(sample) / t_zlib / 1e9}}

    # ---- the reference on a sample: timing + parity ------------------------------------------------------------------------
    from oracle import oracle as O
    if O.have_ref():
        import tempfile
        ref = O.ref()
        m = min(SAMPLE, pairs)
        w = texts[0].size // pairs
        import ctypes
        ref.lib.ref_last_seconds.restype = ctypes.c_double
        want, t_fq = [], 0.0
        for t in texts:
            want.append(ref.fastq_parse(t[:m * w], 3)[0])
            t_fq += ref.lib.ref_last_seconds()  # inside FASTQReader::getNextRead only (not the copies out, not the temporary file)
        got = [L.fastq_parse(t[:m * w], 3)[0] for t in texts]
        same_fq = bool(got[0].same_as(want[0]) and got[1].same_as(want[1]))
        with tempfile.TemporaryDirectory(dir=bench.scratch_dir(len(bases) * 30)) as tmp:
            d = os.path.join(tmp, "idx")
            L.save_index(h, d)
            hc = ref.load_index(d)
            a0, a1 = aln[0][:m].copy(), aln[1][:m].copy()
            sam_ref, _ = ref.sam(hc, want[0], want[1], a0, a1, False, None)
            t_sam = ref.lib.ref_last_seconds()  # Read construction + writePair loop + writer close
            bam_ref, _ = ref.sam(hc, want[0], want[1], a0, a1, False, None, bam=True)
            t_bam = ref.lib.ref_last_seconds()
        sam_got, _ = L.sam(h, got[0], got[1], a0, a1, False, None)
        out["reference_sample"] = {"pairs": m, "threads": 1, "fastq_reads_per_s": 2 * m / t_fq, "sam_reads_per_s": 2 * m / t_sam,
                                   "fastq_arrays_identical": same_fq, "sam_bytes_identical": bool(sam_got == sam_ref),
                                   "sam_bytes": len(sam_ref)}
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_io_fuzz import bam_records
        bam_got, _ = L.sam(h, got[0], got[1], a0, a1, False, None, bam=True)
        out["reference_sample"].update({"bam_reads_per_s": 2 * m / t_bam, "bam_records_identical": bam_records(bytes(bam_got)) == bam_records(bam_ref)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
