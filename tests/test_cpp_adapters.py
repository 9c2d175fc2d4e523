"""The C++ side of the drop-in boundary: the adapters that keep the reference's class names
(include/iftwt_adapters.hpp) and the pipeline example that mirrors main.cpp:91-129 compile
as C++14 and link against libiftwt.so; on a B200 the example runs."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_example(tmp_path):
    from iftwt_rg_b200 import api
    api.load_library()
    exe = str(tmp_path / "main_pipeline")
    libdir = os.path.join(ROOT, "iftwt_rg_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "main_pipeline.cpp"), "-L", libdir, "-liftwt",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_adapters_and_example_compile_and_link(tmp_path):
    exe = build_example(tmp_path)
    assert os.path.exists(exe)
    hdr = open(os.path.join(ROOT, "include", "iftwt_adapters.hpp")).read()
    for name in ("class Graph", "class FlatPoint", "class IFT_PCD", "class Cluster", "getCentroidsCloud",
                 "morph_gradient", "getRegmin", "getRoots", "getAdjacencyList", "getCloud", "pc_to_g"):
        assert name in hdr, name


@pytest.mark.gpu
def test_example_runs_the_reference_pipeline(tmp_path):
    exe = build_example(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    m = re.search(r"points (\d+)\s+vertices (\d+)\s+voxel (\S+)\s+seeds (\d+)\s+regions (\d+)", out.stdout)
    assert m, out.stdout
    assert int(m.group(1)) == 60000 and 0 < int(m.group(2)) <= 60000
    assert int(m.group(4)) >= 1 and 1 <= int(m.group(5)) <= 10


def test_ply_reader_writer(tmp_path):
    """include/iftwt_io.hpp stands in for pcl::io::loadPLYFile / savePLYFile (main.cpp:81, 92, 128-129)."""
    from iftwt_rg_b200 import api
    api.load_library()
    exe = str(tmp_path / "io_test")
    libdir = os.path.join(ROOT, "iftwt_rg_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "io_test.cpp"), "-L", libdir, "-liftwt",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "io ok" in out.stdout, out.stderr


def _build_cpp(tmp_path, name):
    from iftwt_rg_b200 import api
    api.load_library()
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "iftwt_rg_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-L", libdir, "-liftwt",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_graco_adapter_compiles(tmp_path):
    assert os.path.exists(_build_cpp(tmp_path, "graco_test"))


@pytest.mark.gpu
def test_graco_adapter_runs_the_live_main(tmp_path):
    """main.cpp:80-89 (GraCo gc(cloud_t, 5.0); gc.morph_erode) and the Otsu prep through the adapters."""
    out = subprocess.run([_build_cpp(tmp_path, "graco_test")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "graco ok" in out.stdout, out.stdout + out.stderr
