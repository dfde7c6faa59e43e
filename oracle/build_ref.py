#!/usr/bin/env python3
"""Build the UNMODIFIED-ALGORITHM reference (andrewmagis/snap-rnaseq) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

What it does (SURVEY.md section 8c / Appendix B recipe):
  * copies the reference's C++ sources from /root/reference (read-only) to a
    scratch directory under $TMPDIR -- never into this repository;
  * inserts the seven missing `return` statements that g++ >= 8 needs (falling
    off the end of a non-void function is compiled to a trap at -O3); none of
    them touches alignment arithmetic;
  * compiles SNAPLib/*.cpp with -O3 -fPIC and links
        oracle/_ref/libsnapref.so   reference objects + oracle/ref_driver.cpp (C API for ctypes)
        oracle/_ref/snap-rna        the reference CLI (index / transcriptome / single / paired)
        oracle/_ref/ref_unit_tests  the reference's own unit tests (63 KATs)

Outputs go only to oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .so
travels to the GPU box, /root/reference does not).
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (file, 1-based line of the closing brace the function falls off, text that must be on
#  that line, statement to insert before it).  Verified against the reference tree.
PATCHES = [
    ("SNAPLib/Compat.cpp", 931, "}", "return true;"),
    ("SNAPLib/BaseAligner.cpp", 508, "}", "return NotFound;"),
    ("SNAPLib/AlignmentFilter.cpp", 138, "}", "return 1;"),
    ("SNAPLib/GTFReader.cpp", 1300, "}", "return 0;"),
    ("SNAPLib/ContaminationFilter.cpp", 81, "}", "return 0;"),
    ("SNAPLib/ReadReader.cpp", 44, "}", "return false;"),
]
# The two probabilities IntersectingPairedEndAligner::align hands computeMAPQ (probabilityOfAllPairs / probabilityOfBestPair,
# IntersectingPairedEndAligner.cpp:249-252, 741) are locals of align(); they are copied to a thread-local the driver reads so that
# the CUDA path's p_all / p_best are compared with the reference itself and not only with the port.  (file, 1-based line, text that
# must be on that line, line inserted BEFORE it.)  Listed bottom-up so that earlier line numbers stay valid.
INSERT_PATCHES = [
    # INTEGRATION.md section 3, third change: with the alignment core on the GPU the host copy of the hash tables (48 GB at 3.1 Gbp)
    # is dead weight -- the extension needs only the genome text (SAM output) and the seed length.  SNAPB200_GENOME_ONLY (set by
    # snap-rna-b200's main, never by the reference's) makes GenomeIndex::loadFromDirectory skip the overflow table and the hash
    # tables.  Two inserted lines; without the variable the function is unchanged.
    ("SNAPLib/GenomeIndex.cpp", 945, 'snprintf(filenameBuffer,filenameBufferSize,"%s%cGenome",directoryName,PATH_SEP);', "    }"),
    ("SNAPLib/GenomeIndex.cpp", 887, "index->overflowTable = (unsigned *)BigAlloc(index->overflowTableSize * sizeof(*(index->overflowTable)),&index->overflowTableVirtualAllocSize);",
     '    if (getenv("SNAPB200_GENOME_ONLY") != NULL) { index->nHashTables = 0; } else {'),
    ("SNAPLib/IntersectingPairedEndAligner.cpp", 722, "if (bestPairScore == 65536) {",
     "    snapref_pair_p[0] = probabilityOfAllPairs; snapref_pair_p[1] = probabilityOfBestPair;"),
    ("SNAPLib/IntersectingPairedEndAligner.cpp", 141, "void",
     "__thread double snapref_pair_p[2] = {0, 0};"),
]
# operator= without a return value (ContaminationFilter.h:41)
INLINE_PATCHES = [
    ("SNAPLib/ContaminationFilter.h", 41, "count = rhs.count; };", "count = rhs.count; return *this; };"),
    # the one-word change of INTEGRATION.md section 3: CharacterizeSeeds becomes virtual so that the extension's
    # GpuSeedCharacterizer can serve it from device results (no effect on the reference's own behaviour)
    ("SNAPLib/BaseAligner.h", 88, "AlignmentResult", "virtual AlignmentResult"),
    # the other change of INTEGRATION.md section 3: the friend declaration the reference already grants AlignerContext2 in its three
    # context classes, extended to the fusion-interval bookkeeping, so that the extension can append a batch's novel-splice intervals
    # under ONE lock acquisition instead of one GTFReader::IntrachromosomalSplice call (= one acquisition of a process-wide lock with
    # two allocations inside) per candidate.  No effect on the reference's own behaviour.
    ("SNAPLib/GTFReader.h", 56, "class ReadInterval {", "class ReadInterval { friend class AlignerContext2;"),
    ("SNAPLib/GTFReader.h", 134, "class ReadIntervalMap {", "class ReadIntervalMap { friend class AlignerContext2;"),
    ("SNAPLib/GTFReader.h", 318, "class GTFReader {", "class GTFReader { friend class AlignerContext2;"),
]

CXXFLAGS = ["-O3", "-fPIC", "-w", "-fpermissive", "-std=gnu++98", "-Wno-format", "-msse", "-pthread"]


def patch_tree(work):
    for rel, line, expect, stmt in PATCHES:
        p = os.path.join(work, rel)
        lines = open(p, encoding="latin-1").read().split("\n")
        got = lines[line - 1].strip()
        if got != expect:
            raise SystemExit(f"{rel}:{line}: expected {expect!r}, found {got!r} -- reference changed?")
        lines[line - 1] = "    " + stmt + " " + lines[line - 1]
        open(p, "w", encoding="latin-1").write("\n".join(lines))
    for rel, line, expect, stmt in INSERT_PATCHES:
        p = os.path.join(work, rel)
        lines = open(p, encoding="latin-1").read().split("\n")
        got = lines[line - 1].strip()
        if got != expect:
            raise SystemExit(f"{rel}:{line}: expected {expect!r}, found {got!r} -- reference changed?")
        lines.insert(line - 1, stmt)
        open(p, "w", encoding="latin-1").write("\n".join(lines))
    for rel, line, old, new in INLINE_PATCHES:
        p = os.path.join(work, rel)
        lines = open(p, encoding="latin-1").read().split("\n")
        if old not in lines[line - 1]:
            raise SystemExit(f"{rel}:{line}: expected {old!r} -- reference changed?")
        lines[line - 1] = lines[line - 1].replace(old, new)
        open(p, "w", encoding="latin-1").write("\n".join(lines))


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + "\n")
        raise SystemExit(1)
    return r.stdout


def main():
    if not os.path.isdir(os.path.join(REF, "SNAPLib")):
        print(f"build_ref: {REF} not present; keeping whatever is already in {OUT}")
        return 0
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="snapref_")
    try:
        for sub in ("SNAPLib", "apps/snap", "tests"):
            shutil.copytree(os.path.join(REF, sub), os.path.join(work, sub),
                            ignore=shutil.ignore_patterns("datatest", "bin", "*.py", "*.vcxproj*"))
        for root, _, files in os.walk(work):
            os.chmod(root, 0o755)
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)
        patch_tree(work)
        inc = ["-I" + os.path.join(work, "SNAPLib"), "-I" + os.path.join(work, "tests")]
        lib_src = sorted(f for f in os.listdir(os.path.join(work, "SNAPLib")) if f.endswith(".cpp"))
        jobs = [(os.path.join(work, "SNAPLib", f), os.path.join(work, "SNAPLib", f[:-4] + ".o")) for f in lib_src]
        jobs.append((os.path.join(work, "apps/snap/Main.cpp"), os.path.join(work, "Main.o")))
        for f in sorted(os.listdir(os.path.join(work, "tests"))):
            if f.endswith(".cpp"):
                jobs.append((os.path.join(work, "tests", f), os.path.join(work, "tests", f[:-4] + ".o")))
        jobs.append((os.path.join(HERE, "ref_driver.cpp"), os.path.join(work, "ref_driver.o")))

        def cc(job):
            src, obj = job
            run(["g++"] + CXXFLAGS + inc + ["-c", src, "-o", obj])
            return obj

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            objs = list(ex.map(cc, jobs))
        lib_objs = [o for (s, o) in jobs if "/SNAPLib/" in s]
        test_objs = [o for (s, o) in jobs if "/tests/" in s]
        libs = ["-pthread", "-lrt", "-lz"]
        run(["g++", "-shared", "-o", os.path.join(OUT, "libsnapref.so")] + lib_objs +
            [os.path.join(work, "ref_driver.o")] + libs)
        run(["g++", "-o", os.path.join(OUT, "snap-rna")] + lib_objs + [os.path.join(work, "Main.o")] + libs)
        run(["g++", "-o", os.path.join(OUT, "ref_unit_tests")] + lib_objs + test_objs + libs)
        # The drop-in demonstration: the reference's host code + GpuAlignerExtension + libsnapb200.so.  This one IS the
        # product's reference-side binding; it lives in _ref/ only because it links reference objects.
        pkg = os.path.join(os.path.dirname(HERE), "snap_rnaseq_b200")
        so = os.path.join(pkg, "libsnapb200.so")
        if os.path.exists(so):
            run(["g++"] + CXXFLAGS + inc + ["-I" + os.path.join(pkg, "host"), "-I" + os.path.join(os.path.dirname(HERE), "include"),
                 "-c", os.path.join(pkg, "host", "snap_rna_b200_main.cpp"), "-o", os.path.join(work, "b200_main.o")])
            run(["g++", "-o", os.path.join(OUT, "snap-rna-b200")] + lib_objs + [os.path.join(work, "b200_main.o"), so,
                 "-Wl,-rpath," + pkg] + libs)
        print("build_ref: built", ", ".join(sorted(os.listdir(OUT))))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
