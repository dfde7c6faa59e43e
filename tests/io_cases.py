"""Inputs for the FASTQ-parse / SAM-text parity tests, shared by the golden generator, the
host-simulation tests (no GPU) and the CUDA tests.  Everything is derived from seeds; the genome is the one behind
tests/golden/small_index.tar.gz."""
import numpy as np

from snap_rnaseq_b200 import _abi as A
from snap_rnaseq_b200 import synth
from tests_genome import small_genome


def fastq_text(seed=5, n=400, rlen=100, crlf_frac=0.1, partial_tail=True):
    """A FASTQ chunk with the shapes FASTQReader::getNextRead distinguishes: LF and CR LF records, lower-case bases,
    '#' runs at either end of the quality string (Read::clip), reads shorter than 50 after clipping, ragged lengths, ids with
    spaces and /1 suffixes, '@' and '+' as first quality character, and an incomplete record at the end."""
    rng = np.random.default_rng(seed)
    contigs = small_genome()
    sim = synth.simulate(contigs, n, rlen, paired=False, err=0.02, seed=seed, n_rate=0.02)
    b = sim["batches"][0]
    out = []
    for i in range(n):
        s, q = b.read(i)
        L = int(rng.integers(30, rlen + 1)) if rng.random() < 0.3 else rlen
        s, q = s[:L], list(q[:L])
        r = rng.random()
        if r < 0.25:
            k = int(rng.integers(1, 40))
            q[L - k:] = "#" * min(k, L)
        elif r < 0.45:
            k = int(rng.integers(1, 30))
            q[:k] = "#" * min(k, L)
            k2 = int(rng.integers(0, 30))
            if k2:
                q[L - k2:] = "#" * min(k2, L)
        elif r < 0.5:
            q = ["#"] * L
        if rng.random() < 0.1:
            q[0] = "@+"[int(rng.integers(0, 2))]
        q = "".join(q)[:L]
        if rng.random() < 0.2:
            s = s.lower() if rng.random() < 0.5 else s[:L // 2] + s[L // 2:].lower()
        name = f"r{i:x}_{sim['contig'][i]}_{int(sim['start'][i])}"
        if rng.random() < 0.2:
            name += " extra field"
        if rng.random() < 0.3:
            name += "/1"
        nl = "\r\n" if rng.random() < crlf_frac else "\n"
        plus = "+" if rng.random() < 0.7 else "+" + name
        out.append(f"@{name}{nl}{s}{nl}{plus}{nl}{q}{nl}")
    text = "".join(out)
    if partial_tail:
        text += "@partial record\nACGTACGT\n+\n"
    return text.encode()


def _sam_reads(batch, ids, rng, clip=True):
    n = batch.n
    lens = np.diff(batch.offsets).astype(np.int64)
    fc = np.zeros(n, np.uint16)
    cl = lens.astype(np.uint16)
    if clip:
        for i in range(n):
            r = rng.random()
            if r < 0.2 and lens[i] > 70:
                fc[i] = int(rng.integers(1, 10))
                cl[i] = lens[i] - fc[i] - int(rng.integers(0, 10))
            elif r < 0.3 and lens[i] > 70:
                cl[i] = lens[i] - int(rng.integers(1, 15))
    return A.SamReads(batch.offsets, batch.bases, batch.quals, fc, cl, *_ids(ids))


def _ids(ids):
    d, off = A.strings_to_offsets([s.encode() for s in ids])
    return off, d[:off[-1]]


def sam_case(seed=9, n=300, rlen=100, paired=True):
    """Reads + alignments covering SAMFormat::writeRead's branches: both ends mapped (same / different contig, either
    order), one end unmapped, both unmapped, NotFound with a stale location, MAPQ outside 0..70, RC of a read with N and
    with a non-ACGTN byte (SEQ truncated at the NUL COMPLEMENT[] yields), soft clips on either side and strand, wrong
    locations (edit distance > MAX_K-1: CIGAR '*', NM -1), a location that runs off the end of the genome, QNAMEs with
    /1 /2 (and the combinations the reference's condition lets through or not) and with spaces."""
    rng = np.random.default_rng(seed)
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    total = int(piece_off[-1]) + len(list(contigs.values())[-1]) + 500
    sim = synth.simulate(contigs, n, rlen, paired=paired, err=0.02, seed=seed, n_rate=0.02, frag=(max(250, rlen + 20), max(450, rlen + 200)))
    batches = sim["batches"]
    start = sim["start"] if paired else sim["start"][:, None]
    ends = 2 if paired else 1
    ids = [[], []]
    for i in range(n):
        base = f"q{i:x}"
        r = rng.random()
        if r < 0.5:
            a, b = base + "/1", base + "/2"
        elif r < 0.6:
            a, b = base + "/2", base + "/1"
        elif r < 0.65:
            a, b = base + "/1", base + "/1"
        elif r < 0.7:
            a, b = base + "/2", base + "/3"
        elif r < 0.75:
            a, b = base + "/1", base + "x/2"
        elif r < 0.85:
            a, b = base + " comment/1", base + " comment/2"
        else:
            a, b = base, base
        ids[0].append(a)
        ids[1].append(b)
    # a few reads get a byte COMPLEMENT[] maps to NUL, so that an RC alignment truncates SEQ
    for e in range(ends):
        bb = batches[e].bases
        for i in rng.choice(n, size=max(1, n // 40), replace=False):
            bb[int(batches[e].offsets[i]) + int(rng.integers(0, rlen))] = ord("R")
    reads = [_sam_reads(batches[e], ids[e], rng) for e in range(ends)]
    aln = [np.zeros(n, A.SAM_ALIGNMENT) for _ in range(ends)]
    for i in range(n):
        c = int(sim["contig"][i])
        sw = int(sim["strand"][i]) == 1
        for e in range(ends):
            a = aln[e][i]
            if paired:
                left = (e == 0) != sw
                pos = int(start[i][0 if left else 1])
                direction = A.FORWARD if left else A.RC
            else:
                pos = int(start[i][0])
                direction = A.RC if sw else A.FORWARD
            fc = int(reads[e].front_clip[i])
            full = int(reads[e].offsets[i + 1] - reads[e].offsets[i])
            cl = int(reads[e].clipped_len[i])
            # the aligner sees the clipped read: its location is where that piece starts on the genome
            shift = fc if direction == A.FORWARD else full - cl - fc
            a["location"] = int(piece_off[c]) + pos + shift
            a["direction"] = direction
            a["status"] = A.SINGLE_HIT if rng.random() < 0.8 else A.MULTIPLE_HITS
            a["mapq"] = int(rng.integers(0, 71))
            r = rng.random()
            if r < 0.10:
                a["status"] = A.NOT_FOUND  # with a stale location and direction left in place
            elif r < 0.15:
                a["status"], a["location"] = A.NOT_FOUND, A.INVALID_LOCATION
            elif r < 0.20:
                a["location"] = int(rng.integers(500, total - 600))  # wrong place: LV fails
            elif r < 0.23:
                a["location"] = total - int(rng.integers(1, rlen // 2))  # runs off the end of the genome
            elif r < 0.27:
                a["mapq"] = int(rng.choice([-5, 71, 200]))
            elif r < 0.32:
                a["location"] = int(piece_off[(c + 1) % len(piece_off)]) + int(rng.integers(0, 2000))  # other contig
    return reads, aln


def clipped_for_cigar(reads, aln):
    """What computeCigarString aligns: the clipped read at its location (InvalidGenomeLocation when NotFound)."""
    b = reads.clipped_batch()
    loc = np.where(aln["status"] == A.NOT_FOUND, A.INVALID_LOCATION, aln["location"]).astype(np.uint32)
    return b, loc, aln["direction"].copy()
