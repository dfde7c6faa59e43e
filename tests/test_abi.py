"""The C-ABI library loads and exports every symbol include/snapb200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "snapb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snapb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("snapb200_index_open", "snapb200_single_batch", "snapb200_single_multihit_batch", "snapb200_paired_batch",
              "snapb200_cigar_batch", "snapb200_stats_get", "snapb200_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import snap_rnaseq_b200 as S
    from snap_rnaseq_b200 import build as B
    B.build()
    lib = C.CDLL(S.SO_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/snapb200.h but not exported"
    lib.snapb200_abi_version.restype = C.c_int
    assert lib.snapb200_abi_version() == 3


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof of every struct, as gcc sees include/snapb200.h, against the ctypes/numpy mirrors."""
    import subprocess

    from snap_rnaseq_b200 import _abi as A
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "snapb200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(snapb200_single_params), sizeof(snapb200_paired_params),'
        ' sizeof(snapb200_single_result), sizeof(snapb200_paired_result), sizeof(snapb200_index_info), sizeof(snapb200_stats),'
        ' offsetof(snapb200_single_result,p_all), offsetof(snapb200_paired_result,p_all), offsetof(snapb200_paired_result,n_lv_calls));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert got[0] == C.sizeof(A.SingleParams)
    assert got[1] == C.sizeof(A.PairedParams)
    assert got[2] == A.SINGLE_RESULT.itemsize
    assert got[3] == A.PAIRED_RESULT.itemsize
    assert got[4] == C.sizeof(A.IndexInfo)
    assert got[5] == A.STATS_WORDS * 8
    assert got[6] == A.SINGLE_RESULT.fields["p_all"][1]
    assert got[7] == A.PAIRED_RESULT.fields["p_all"][1]
    assert got[8] == A.PAIRED_RESULT.fields["n_lv_calls"][1]


def test_no_gpu_means_failure_not_fallback():
    """Without a CUDA device every compute entry point must fail loudly (there is no CPU path)."""
    import snap_rnaseq_b200 as S
    L = S.lib()
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        L.lv(1, [b"abc"], [b"abc"], None, [2])
    with pytest.raises(RuntimeError):
        L.load_index("/nonexistent")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        L.fastq_parse(b"@a\nACGT\n+\nIIII\n", 0)


def test_io_struct_layouts_match_header(tmp_path):
    """snapb200_sam_reads / snapb200_sam_alignment as gcc sees them against the ctypes / numpy mirrors."""
    import subprocess

    from snap_rnaseq_b200 import _abi as A
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "snapb200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(snapb200_sam_reads), offsetof(snapb200_sam_reads,front_clip),'
        ' offsetof(snapb200_sam_reads,ids), sizeof(snapb200_sam_alignment), offsetof(snapb200_sam_alignment,mapq),'
        ' offsetof(snapb200_sam_alignment,status), offsetof(snapb200_sam_alignment,skip));'
        'printf("%zu %zu %zu %zu %zu\\n", sizeof(snapb200_filter_result), offsetof(snapb200_filter_result,status), sizeof(snapb200_filter_event),'
        ' offsetof(snapb200_filter_event,pos), sizeof(snapb200_filter_params));'
        'printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(snapb200_filter_result,aligned_as_pair), sizeof(snapb200_rna_params), offsetof(snapb200_rna_params,filter),'
        ' sizeof(snapb200_rna_view), offsetof(snapb200_rna_view,seg_offsets), offsetof(snapb200_rna_view,device_ms));'
        'printf("%zu %zu\\n", offsetof(snapb200_sam_alignment,is_transcriptome), offsetof(snapb200_sam_alignment,tlocation));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    assert got[0] == C.sizeof(A.SamReadsStruct)
    assert got[1] == A.SamReadsStruct.front_clip.offset and got[2] == A.SamReadsStruct.ids.offset
    assert got[3] == A.SAM_ALIGNMENT.itemsize
    assert got[4] == A.SAM_ALIGNMENT.fields["mapq"][1] and got[5] == A.SAM_ALIGNMENT.fields["status"][1]
    assert got[6] == A.SAM_ALIGNMENT.fields["skip"][1]
    assert got[-2] == A.SAM_ALIGNMENT.fields["is_transcriptome"][1] and got[-1] == A.SAM_ALIGNMENT.fields["tlocation"][1]
    assert got[7] == A.FILTER_RESULT.itemsize and got[8] == A.FILTER_RESULT.fields["status"][1]
    assert got[9] == A.FILTER_EVENT.itemsize and got[10] == A.FILTER_EVENT.fields["pos"][1] and got[11] == C.sizeof(A.FilterParams)
    assert got[12] == A.FILTER_RESULT.fields["aligned_as_pair"][1]
    assert got[13] == C.sizeof(A.RnaParams) and got[14] == A.RnaParams.filter.offset
    assert got[15] == C.sizeof(A.RnaView) and got[16] == A.RnaView.seg_offsets.offset and got[17] == A.RnaView.device_ms.offset


def test_product_does_not_reference_the_oracle():
    """The package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "snap_rnaseq_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "libsnapref" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
