"""Synthetic inputs for the alignment hot path (there is no network: SURVEY.md section 8d).

The reference's WGsim.{h,cpp} only parses WGsim-style read names (SNAPLib/WGsim.cpp:40-174); it is not a
simulator, so this module is one: uniform-random and repeat-injected genomes, and reads under the WGsim
model (uniform start, random strand, per-base substitution rate, a fraction of mutations being short
indels, FR pairs with a uniform fragment size), named `<contig>_<start1>_<end1>_...` like wgsim does so
that the reference's `-e` column keeps working.  Everything is seeded and vectorised with numpy.
"""
import numpy as np

from ._abi import Batch

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTNn", b"TGCANn"):
    _COMP[_a] = _b

# Phred+33 quality alphabet: eight levels, >= 95 % of draws are >= Q20 so the default quality gate
# (-fm 20 -fp 90, SNAPLib/Read.h:422-433) passes and lv_phredToProbability sees several values.
QUAL_LEVELS = np.frombuffer(b"IGDA>;5+", dtype=np.uint8)  # Q40 38 35 32 29 26 20 10
QUAL_P = np.array([0.40, 0.20, 0.12, 0.10, 0.07, 0.04, 0.04, 0.03])


def random_contigs(lengths, seed=20, prefix="chr"):
    """i.i.d. uniform ACGT contigs: dict name -> uint8 array, in insertion order."""
    rng = np.random.default_rng(seed)
    return {f"{prefix}{i + 1}": BASES[rng.integers(0, 4, size=int(n), dtype=np.uint8)] for i, n in enumerate(lengths)}


def inject_repeats(contigs, frac=0.05, seed=21, min_len=200, max_len=5000, max_copies=2000, max_div=0.03):
    """Copy segments around the genome until ~`frac` of the bases sit in repeats (copy number 2..max_copies,
    0..max_div per-copy divergence), so that seeds with 2, >300 and >16000 hits exist."""
    rng = np.random.default_rng(seed)
    names = list(contigs)
    total = sum(len(contigs[n]) for n in names)
    budget = int(total * frac)
    while budget > 0:
        src_c = contigs[names[rng.integers(len(names))]]
        ln = int(rng.integers(min_len, max_len + 1))
        if ln >= len(src_c) // 4:
            ln = max(50, len(src_c) // 8)
        s = int(rng.integers(0, len(src_c) - ln))
        seg = src_c[s:s + ln].copy()
        # copy number: log-uniform so that a few families are huge
        copies = int(np.exp(rng.uniform(np.log(2), np.log(max_copies))))
        copies = max(1, min(copies, budget // ln + 1))
        div = rng.uniform(0, max_div)
        for _ in range(copies):
            dst = contigs[names[rng.integers(len(names))]]
            if len(dst) <= ln:
                continue
            d = int(rng.integers(0, len(dst) - ln))
            c = seg.copy()
            nm = rng.binomial(ln, div)
            if nm:
                pos = rng.integers(0, ln, size=nm)
                c[pos] = BASES[rng.integers(0, 4, size=nm)]
            dst[d:d + ln] = c
        budget -= ln * copies
    return contigs


def write_fasta(path, contigs, width=100):
    with open(path, "wb") as f:
        for name, seq in contigs.items():
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def snap_layout(contigs, padding=500):
    """The reference's in-memory genome: padding 'n' before every contig and after the last one
    (SNAPLib/FASTA.cpp:68-126).  Returns (bases uint8, piece_offsets uint32)."""
    parts, offs, pos = [], [], 0
    pad = np.full(padding, ord("n"), np.uint8)
    for seq in contigs.values():
        parts.append(pad)
        pos += padding
        offs.append(pos)
        parts.append(seq)
        pos += len(seq)
    parts.append(pad)
    return np.concatenate(parts), np.array(offs, np.uint32)


def _mutate(frag, rlen, rng, err, indel_frac, n_rate):
    """frag: [n, rlen+8] template windows -> [n, rlen] reads with substitutions, at most one short indel
    per read, and a few Ns."""
    n = frag.shape[0]
    idx = np.broadcast_to(np.arange(rlen, dtype=np.int64), (n, rlen)).copy()
    n_mut = rng.binomial(rlen, err, size=n)
    has_indel = rng.random(n) < 1.0 - (1.0 - indel_frac) ** n_mut
    ilen = np.minimum(rng.geometric(0.7, size=n), 4)
    is_ins = rng.random(n) < 0.5
    ipos = rng.integers(5, rlen - 5, size=n)
    col = np.arange(rlen)[None, :]
    dele = has_indel & ~is_ins
    idx += np.where(dele[:, None] & (col >= ipos[:, None]), ilen[:, None], 0)
    ins = has_indel & is_ins
    idx -= np.where(ins[:, None] & (col >= (ipos + ilen)[:, None]), ilen[:, None], 0)
    reads = np.take_along_axis(frag, idx, axis=1)
    ins_cols = ins[:, None] & (col >= ipos[:, None]) & (col < (ipos + ilen)[:, None])
    k = int(ins_cols.sum())
    if k:
        reads[ins_cols] = BASES[rng.integers(0, 4, size=k)]
    sub = rng.random((n, rlen)) < err * (1.0 - indel_frac)
    k = int(sub.sum())
    if k:
        reads[sub] = BASES[(np.searchsorted(BASES, reads[sub]) + rng.integers(1, 4, size=k)) % 4]
    if n_rate > 0:
        which = np.nonzero(rng.random(n) < n_rate)[0]
        for r in which:
            reads[r, rng.integers(0, rlen, size=int(rng.integers(1, 4)))] = ord("N")
    return reads


def _quals(n, rlen, rng):
    return QUAL_LEVELS[rng.choice(len(QUAL_LEVELS), size=(n, rlen), p=QUAL_P)]


def _revcomp(a):
    return _COMP[a[:, ::-1]]


def _to_batch(reads, quals):
    n, rlen = reads.shape
    off = (np.arange(n + 1, dtype=np.uint64) * rlen).astype(np.uint32)
    return Batch(reads.reshape(-1), quals.reshape(-1), off)


def simulate(contigs, n, read_len, *, paired=False, err=0.02, indel_frac=0.15, n_rate=0.005, frag=(250, 450),
             seed=7, junk_frac=0.0):
    """Simulate n reads (or n FR pairs).  Returns dict(batches=[Batch] or [Batch, Batch], contig=idx array,
    start=[n] or [n,2] 0-based template starts, strand=[n], names=list of contig names)."""
    rng = np.random.default_rng(seed)
    names = list(contigs)
    seqs = [contigs[k] for k in names]
    lens = np.array([len(s) for s in seqs], np.int64)
    span_need = (frag[1] if paired else read_len) + 16
    ok = lens > span_need + 16
    w = np.where(ok, lens, 0).astype(np.float64)
    cidx = rng.choice(len(seqs), size=n, p=w / w.sum())
    flen = rng.integers(frag[0], frag[1] + 1, size=n) if paired else np.full(n, read_len)
    flen = np.maximum(flen, read_len)
    start = (rng.random(n) * (lens[cidx] - flen - 16)).astype(np.int64)
    strand = rng.integers(0, 2, size=n)
    win = read_len + 8
    col = np.arange(win)[None, :]
    # gather template windows per contig (vectorised inside each contig)
    left = np.empty((n, win), np.uint8)
    right = np.empty((n, win), np.uint8) if paired else None
    for c in range(len(seqs)):
        m = np.nonzero(cidx == c)[0]
        if m.size == 0:
            continue
        s = seqs[c]
        left[m] = s[start[m, None] + col]
        if paired:
            # right mate: the last `win` bases of the fragment (+8 slack to the left), reverse-complemented
            rs = start[m] + flen[m] - read_len
            right[m] = _COMP[s[(rs[:, None] + read_len - 1) - col]]
    if not paired:
        reads = _mutate(left, read_len, rng, err, indel_frac, n_rate)
        rc = strand == 1
        reads[rc] = _revcomp(reads[rc])
        if junk_frac > 0:
            j = rng.random(n) < junk_frac
            reads[j] = BASES[rng.integers(0, 4, size=(int(j.sum()), read_len))]
        quals = _quals(n, read_len, rng)
        return dict(batches=[_to_batch(reads, quals)], contig=cidx, start=start, strand=strand, names=names)
    r_left = _mutate(left, read_len, rng, err, indel_frac, n_rate)
    r_right = _mutate(right, read_len, rng, err, indel_frac, n_rate)
    # strand 0: mate1 = forward left end, mate2 = reverse-complemented right end; strand 1: swapped
    sw = strand == 1
    m1 = np.where(sw[:, None], r_right, r_left)
    m2 = np.where(sw[:, None], r_left, r_right)
    if junk_frac > 0:
        j = rng.random(n) < junk_frac
        m2[j] = BASES[rng.integers(0, 4, size=(int(j.sum()), read_len))]
    q1, q2 = _quals(n, read_len, rng), _quals(n, read_len, rng)
    st2 = np.stack([start, start + flen - read_len], axis=1)
    return dict(batches=[_to_batch(m1, q1), _to_batch(m2, q2)], contig=cidx, start=st2, strand=strand, names=names)


def write_fastq(path, batch, sim, mate=0):
    """WGsim-style names: <contig>_<start1>_<end1>_0:0:0_0:0:0_<serial>/<mate> (1-based coordinates)."""
    names = sim["names"]
    with open(path, "wb") as f:
        for i in range(batch.n):
            b, q = batch.read(i)
            st = sim["start"][i]
            s1 = int(st[0] if np.ndim(st) else st) + 1
            e1 = int((st[1] if np.ndim(st) else st)) + len(b)
            f.write(f"@{names[sim['contig'][i]]}_{s1}_{e1}_0:0:0_0:0:0_{i:x}/{mate + 1}\n{b}\n+\n{q}\n".encode())


def make_gtf(path, contigs, seed=31, gene_frac=0.03, decoy="chrDecoy"):
    """Synthetic annotation: a decoy transcript first (its id sorts lowest; the reference dereferences a NULL piece for
    transcriptome hits that start in the padding before the FIRST transcript), then non-overlapping
    multi-exon '+'-strand genes (2-6 exons of 150-600 bp, introns 200-2000 bp) over ~gene_frac of each contig."""
    rng = np.random.default_rng(seed)
    lines = []
    attr = 'gene_id "{g}"; transcript_id "{t}"; gene_name "{g}"; transcript_name "{t}";'
    if decoy in contigs:
        lines.append("\t".join([decoy, "synth", "exon", "101", "900", ".", "+", ".", attr.format(g="AAAA_decoy", t="AAAA_decoy_t")]))
    gid = 0
    for name, seq in contigs.items():
        if name == decoy:
            continue
        pos, budget = 2000, int(len(seq) * gene_frac)
        while budget > 0 and pos < len(seq) - 12000:
            n_ex = int(rng.integers(2, 7))
            g, t = f"G{gid:05d}", f"T{gid:05d}"
            p = pos
            for _ in range(n_ex):
                ln = int(rng.integers(150, 601))
                lines.append("\t".join([name, "synth", "exon", str(p + 1), str(p + ln), ".", "+", ".", attr.format(g=g, t=t)]))
                budget -= ln
                p += ln + int(rng.integers(200, 2001))
            gid += 1
            pos = p + int(rng.integers(3000, 30000))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(lines)


def transcripts_from_gtf(gtf_path, contigs):
    """Transcript sequences (exons concatenated in file order, '+' strand) of a GTF written by make_gtf: {transcript_id: bases}."""
    import re
    exons = {}
    for line in open(gtf_path):
        f = line.rstrip("\n").split("\t")
        if len(f) < 9 or f[2] != "exon":
            continue
        tid = re.search(r'transcript_id "([^"]+)"', f[8]).group(1)
        exons.setdefault(tid, []).append((f[0], int(f[3]) - 1, int(f[4])))
    return {t: np.concatenate([contigs[c][a:b] for c, a, b in ex]) for t, ex in exons.items()}


def simulate_rna(contigs, gtf_path, n, read_len, *, spliced_frac=0.5, chimeric_frac=0.01, err=0.02, seed=11, junk_frac=0.005,
                 decoy="chrDecoy"):
    """BASELINE.json configs[3]: n FR pairs, `spliced_frac` of the fragments drawn from spliced transcripts (so reads span exon
    junctions), the rest from the genome, `chimeric_frac` of the pairs with mate 2 taken from another pair.  Returns [Batch, Batch]."""
    rng = np.random.default_rng(seed)
    real = {k: v for k, v in contigs.items() if k != decoy}
    tx = {k: v for k, v in transcripts_from_gtf(gtf_path, contigs).items() if "decoy" not in k}
    frag = (max(250, read_len + 20), max(450, read_len + 200))
    tx = {k: v for k, v in tx.items() if len(v) > frag[1] + 64}
    n_sp = int(n * spliced_frac) if tx else 0
    parts = [simulate(real, n - n_sp, read_len, paired=True, err=err, seed=seed + 1, junk_frac=junk_frac, frag=frag)["batches"]]
    if n_sp:
        parts.append(simulate(tx, n_sp, read_len, paired=True, err=err, seed=seed + 2, frag=frag)["batches"])
    order = rng.permutation(n)
    out = []
    for mate in range(2):
        reads = np.concatenate([p[mate].bases.reshape(-1, read_len) for p in parts])[order]
        quals = np.concatenate([p[mate].quals.reshape(-1, read_len) for p in parts])[order]
        out.append([reads, quals])
    k = int(n * chimeric_frac)
    if k:
        a, b = rng.choice(n, size=k, replace=False), rng.choice(n, size=k, replace=False)
        out[1][0][a], out[1][1][a] = out[1][0][b].copy(), out[1][1][b].copy()
    return [_to_batch(r, q) for r, q in out]


def write_fastq_plain(path, batch, mate=0, prefix="r"):
    """FASTQ with ids <prefix><serial>/<mate> (mates of a pair share the id up to the slash, as Read::checkIdMatch wants)."""
    with open(path, "wb") as f:
        for i in range(batch.n):
            b, q = batch.read(i)
            f.write(f"@{prefix}{i:x}/{mate + 1}\n{b}\n+\n{q}\n".encode())


def fastq_fixed(batch, mate=0):
    """FASTQ text (uint8 array) of a batch of equal-length reads, built column-wise: @r<8 hex digits>/<mate+1> LF bases LF + LF
    qualities LF.  Used by the I/O-edge measurements (scripts/io_bench.py, bench.py io_edges)."""
    n = batch.n
    rlen = int(batch.offsets[1] - batch.offsets[0]) if n else 0
    assert n == 0 or int(batch.offsets[-1]) == n * rlen
    w = 1 + 11 + 1 + rlen + 3 + rlen + 1
    rec = np.empty((n, w), np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    idx = np.arange(n, dtype=np.uint64)
    hexd = np.frombuffer(b"0123456789abcdef", np.uint8)
    for k in range(8):
        rec[:, 2 + k] = hexd[(idx >> np.uint64(4 * (7 - k))) & np.uint64(15)]
    rec[:, 10] = ord("/")
    rec[:, 11] = ord("1") + mate
    rec[:, 12] = 10
    rec[:, 13:13 + rlen] = batch.bases.reshape(n, rlen)
    rec[:, 13 + rlen] = 10
    rec[:, 14 + rlen] = ord("+")
    rec[:, 15 + rlen] = 10
    rec[:, 16 + rlen:16 + 2 * rlen] = batch.quals.reshape(n, rlen)
    rec[:, 16 + 2 * rlen] = 10
    return rec.reshape(-1)
