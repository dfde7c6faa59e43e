"""BGZF blocks (the compressed container of BAM output; csrc/bgzf.h, csrc/bgzf_kernels.cuh).
The reference compresses every 64 KiB chunk with zlib inside a gzip member that carries the "BC" field
(SNAPLib/GzipDataWriter.cpp:281-340).  What is checked here is what a BGZF reader needs: every member is a well-formed BGZF block
(magic, BC field, BSIZE = the member's size, ISIZE), the members inflate -- with zlib itself -- to exactly the input, and the CRC-32
words are right (zlib verifies them).  The bytes of the deflate streams are this library's own and are only compared between its two
implementations: the serial specification (host simulation) and the CUDA kernels must agree byte for byte."""
import gzip
import zlib

import numpy as np
import pytest

from test_io_edges import hostsim  # noqa: F401  (fixture)

CHUNK = 65024


def cases():
    rng = np.random.default_rng(77)
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    out = {
        "empty": b"",
        "one byte": b"a",
        "one symbol": b"\0" * 70000,
        "incompressible": bytes(rng.integers(0, 256, 200000, dtype=np.uint8)),             # stored blocks
        "bases": bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), 300001, p=[.25, .25, .25, .24, .01])),
        "skewed": bytes(np.minimum(255, rng.geometric(0.3, 150000)).astype(np.uint8)),
        "depth over 15": b"".join(bytes([i]) * fib[i] for i in range(24)),                   # Fibonacci weights: the tree must be flattened
        "exact chunks": bytes(rng.integers(0, 4, 2 * CHUNK, dtype=np.uint8)),
        "chunk plus one": bytes(rng.integers(0, 4, CHUNK + 1, dtype=np.uint8)),
    }
    # what the writer really sees: BAM-like records (binary head, name, nibbles, qualities from a small alphabet, tags)
    recs = []
    for i in range(4000):
        head = rng.integers(0, 256, 36, dtype=np.uint8).tobytes()
        name = b"read%07x\0" % i
        seq = rng.integers(0, 256, 50, dtype=np.uint8).tobytes()
        qual = bytes(rng.choice(np.array([2, 11, 25, 30, 37, 40], np.uint8), 100))
        recs.append(head + name + seq + qual + b"PGZSNAP\0NMi" + bytes([i % 5, 0, 0, 0]))
    out["bam-like"] = b"".join(recs)
    return out


def check_container(z, data, chunk):
    assert gzip.decompress(z) == data  # zlib inflates every member and verifies CRC-32 and ISIZE
    p, n_blocks, covered = 0, 0, 0
    while p < len(z):
        assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 10:p + 12] == b"\x06\x00" and z[p + 12:p + 16] == b"BC\x02\x00"
        size = int.from_bytes(z[p + 16:p + 18], "little") + 1
        isize = int.from_bytes(z[p + size - 4:p + size], "little")
        assert isize <= chunk and size <= 65536
        assert int.from_bytes(z[p + size - 8:p + size - 4], "little") == zlib.crc32(data[covered:covered + isize])
        covered += isize
        p += size
        n_blocks += 1
    assert p == len(z) and covered == len(data) and n_blocks == max(1, (len(data) + chunk - 1) // chunk)
    return n_blocks


def test_bgzf_serial_specification_against_zlib(hostsim):
    ratios = {}
    for name, data in cases().items():
        for chunk in (CHUNK, 1000) if len(data) < 250000 else (CHUNK,):
            z, _ = hostsim.bgzf_compress(data, chunk)
            check_container(z, data, chunk)
            if chunk == CHUNK and data:
                ratios[name] = len(z) / len(data)
    assert ratios["incompressible"] < 1.001 and ratios["bases"] < 0.30 and ratios["one symbol"] < 0.14 and ratios["bam-like"] < 0.85, ratios
    # the slice-wise CRC the kernel uses (32 registers combined through the zero-shift matrices) is the CRC
    import ctypes as C
    for name, data in cases().items():
        if not data:
            continue
        a = np.frombuffer(data[:CHUNK], np.uint8)
        for slices in (1, 2, 7, 32):
            got = hostsim.lib.hostsim_bgzf_crc_sliced(a.ctypes.data_as(C.c_void_p), C.c_uint32(a.size), C.c_uint32(slices)) & 0xffffffff
            assert got == zlib.crc32(data[:CHUNK]), (name, slices)


@pytest.mark.gpu
def test_bgzf_cuda_matches_the_serial_specification_and_inflates(cuda, hostsim):
    for name, data in cases().items():
        for chunk in (0, 1000) if len(data) < 250000 else (0,):
            z, off = cuda.bgzf_compress(data, chunk)
            n_blocks = check_container(z, data, chunk or CHUNK)
            want, _ = hostsim.bgzf_compress(data, chunk or CHUNK)
            assert z == want, (name, chunk, len(z), len(want))
            assert len(off) == n_blocks + 1 and off[0] == 0 and int(off[-1]) == len(z)
    # a stream of the size a writer buffer has: 48 MB of BAM-like bytes, 775 blocks
    big = cases()["bam-like"] * 60
    z, off = cuda.bgzf_compress(big)
    assert gzip.decompress(z) == big and len(z) < 0.85 * len(big)
    assert cuda.bgzf_last_kernel_ms() > 0
    with pytest.raises(RuntimeError, match="chunk"):
        cuda.bgzf_compress(b"abc", 70000)
