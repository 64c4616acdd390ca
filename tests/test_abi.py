"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/fwgpu.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

import fwload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fw():
    import __graft_entry__ as ge
    ge.build()
    return fwload.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fwgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fw_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(fw):
    L = fw.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for s in declared:
        assert hasattr(L, s), "libfwgpu.so does not export %s" % s
    assert sorted(fw.ABI_SYMBOLS) == declared


def test_testresult_layout(fw):
    # src/types.jl:140-145: isbits {Float64, Float64, Int64, Bool} = 32 bytes with padding
    assert C.sizeof(fw.TestResult) == 32
    assert fw.TestResult.pval.offset == 8 and fw.TestResult.df.offset == 16 and fw.TestResult.suff_power.offset == 24


def test_no_cpu_fallback(fw):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fw.FwError) as ei:
        fw.Engine(0)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under flashweave.jl_b200/ may import, link or dlopen it."""
    pkg = os.path.join(ROOT, "flashweave.jl_b200")
    bad = re.compile(r"from\s+oracle|import\s+oracle|liboracle|\bfwo\b|#include\s+\"[^\"]*oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f


def test_host_graph_assembly(fw):
    import numpy as np
    # two targets, one edge seen from both sides with different weights -> larger |w| wins (misc.jl:201-218)
    res = fw.HitonResult(np.array([0, 1]), np.array([0, 1, 2]), np.array([1, 1]), np.array([1, 0]), np.array([0.2, 0.5]), np.array([1e-3, 1e-4]),
                         np.array([0, 0]), None, None, None, np.array([3, 4]), 7)
    uni = fw.NbrCSR(np.array([0, 1, 2]), np.array([1, 0]), np.array([0.3, 0.3]), np.array([1e-5, 1e-5]))
    assert fw.assemble_graph(res, uni, "fz") == [(0, 1, 0.5)]
    assert list(fw.target_order(fw.NbrCSR(np.array([0, 2, 2, 3]), None, None, None))) == [1, 2, 0]
    assert list(fw.shard_targets(np.arange(10), 1, 4)) == [1, 5, 9]


def test_edgelist_format(fw, tmp_path):
    """write_edgelist follows src/io.jl:338-359: `# header`, `# meta mask`, then `name<TAB>name<TAB>weight` with Julia's
    shortest round-trip float text (Python's repr gives the same digits)."""
    path = tmp_path / "net.edgelist"
    fw.write_edgelist(str(path), [(3, 6, 0.1891748160123825), (3, 13, -0.31248563528060913)], p=14)
    lines = open(path).read().split("\n")
    assert lines[0] == "# header\t" + ",".join("X%d" % i for i in range(1, 15))
    assert lines[1] == "# meta mask\t" + ",".join(["false"] * 14)
    assert lines[2] == "X4\tX7\t0.1891748160123825" and lines[3] == "X4\tX14\t-0.31248563528060913" and lines[4] == ""
    fw.write_edgelist(str(path), [(0, 1, 1.0)], header=["a", "b", "m"], meta_mask=[False, False, True])
    assert open(path).read() == "# header\ta,b,m\n# meta mask\tfalse,false,true\na\tb\t1.0\n"


def test_edgelist_writer_reproduces_reference_files_byte_for_byte(fw, tmp_path, golden_dir):
    """All 8 expected graphs of test/data/learning_expected (verbatim copies under tests/golden/edgelists): parsing a file and
    writing it back with write_edgelist gives the identical bytes - header lines, edge order, Julia's float text."""
    d = os.path.join(golden_dir, "edgelists")
    files = sorted(os.listdir(d))
    assert len(files) == 8
    for fn in files:
        raw = open(os.path.join(d, fn)).read()
        lines = raw.split("\n")
        header = lines[0].split("\t")[1].split(",")
        mask = [m == "true" for m in lines[1].split("\t")[1].split(",")]
        pos = {h: i for i, h in enumerate(header)}
        edges = []
        for ln in lines[2:]:
            if ln:
                a, b, w = ln.split("\t")
                edges.append((pos[a], pos[b], float(w)))
        assert all(a < b for a, b, _ in edges)
        out = tmp_path / fn
        fw.write_edgelist(str(out), sorted(edges), header=header, meta_mask=mask)             # sorted(): the order assemble_graph produces
        assert open(out).read() == raw, fn
    for x, want in [(1e-5, "1.0e-5"), (5e-5, "5.0e-5"), (1e-4, "0.0001"), (999999.0, "999999.0"), (1e6, "1.0e6"), (1234567.0, "1.234567e6"),
                    (float("nan"), "NaN"), (-0.0, "-0.0"), (1.2345e-7, "1.2345e-7"), (0.30000001192092896, "0.30000001192092896")]:
        assert fw.julia_float_str(x) == want, (x, fw.julia_float_str(x))


def test_c_host_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """include/fwgpu.h is plain C99; examples/fw_demo.c (the ccall sequence of INTEGRATION.md written in C) compiles with
    -pedantic, links against libfwgpu.so, and either runs (GPU present) or stops at fw_create with the no-fallback message."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "fw_demo")
    libdir = os.path.join(root, "flashweave.jl_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "examples", "fw_demo.c"), "-L" + libdir, "-lfwgpu", "-Wl,-rpath," + libdir, "-lm", "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    if r.returncode == 0:
        assert "conditional tests" in r.stdout
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stdout, r.stdout
