"""CPU-side checks of the drop-in boundary: libiftwt.so loads, exports every symbol
include/iftwt.h declares, has no torch types in its signatures, and fails loudly
(no CPU fallback) when no B200 is visible."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "iftwt.h")).read()
    return re.findall(r"IFTWT_API\s+[\w\s\*]+?\b(iftwt_\w+)\s*\(", src)


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("iftwt_create", "iftwt_destroy", "iftwt_set_cloud", "iftwt_knn", "iftwt_knn_query",
                 "iftwt_knn_graph", "iftwt_normals", "iftwt_region_seeds", "iftwt_ift", "iftwt_segment",
                 "iftwt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from iftwt_rg_b200 import api
    lib = api.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.iftwt_version() == 100


def test_header_is_plain_c_and_compiles_with_gcc(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "iftwt.h"\nint main(void){ int (*f)(void) = iftwt_version; return f == 0; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])
    text = open(os.path.join(ROOT, "include", "iftwt.h")).read()
    assert "torch" not in text and "at::" not in text


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing in the product may import, include,
    link or dlopen it (comments that cite it for parity are fine)."""
    pkg = os.path.join(ROOT, "iftwt_rg_b200")
    bad = re.compile(r"(from\s+oracle|import\s+oracle|pyoracle|liboracle|#\s*include\s*[\"<][^\">]*oracle|"
                     r"CDLL\([^)]*oracle|-loracle)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not bad.search(txt), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert not bad.search(open(os.path.join(ROOT, "include", f)).read()), f
    # neither do the example, the tools or the reference checker's third-party stand-ins (pyref included)
    for d in ("examples", "tools"):
        for f in os.listdir(os.path.join(ROOT, d)):
            if f.endswith((".py", ".cpp", ".hpp", ".h")):
                assert not bad.search(open(os.path.join(ROOT, d, f), errors="ignore").read()), f
    bad_ref = re.compile(r"(from\s+oracle|import\s+oracle|pyref|libref_|oracle/_ref|oracle/shim)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                assert not bad_ref.search(open(os.path.join(dp, f), errors="ignore").read()), f


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from iftwt_rg_b200 import api
    with pytest.raises(api.IftwtError) as e:
        api.Context(0)
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)
    # the batch and the multi-GPU entry points fail the same way: no device, no fallback
    for make in (lambda: api.Batch(0, 2), lambda: api.Comm.all([0, 0]), lambda: api.Comm.all([0])):
        with pytest.raises(api.IftwtError) as e:
            make()
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)
