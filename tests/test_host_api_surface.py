"""The drop-in boundary as a compile check: tests/host_api_surface.cpp names every member of the reference's
public C++ class surface on the hot path with its exact type.  It must compile against the veritas_b200 host classes — and,
where /root/reference is present (the build container), against the unmodified reference headers, which proves that the list
itself is the reference's.  Syntax check only: nothing is linked or run, no device needed."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_api_surface.cpp")
REF = "/root/reference"


def syntax_only(flags):
    r = subprocess.run(["g++", "-std=c++14", "-fopenmp", "-fsyntax-only", "-w"] + flags + [SRC], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stderr


def test_host_classes_offer_the_reference_class_surface():
    rc, err = syntax_only(["-DVRT_HOST_BUILD", "-I" + os.path.join(ROOT, "veritas_b200", "host")])
    assert rc == 0, err[:4000]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_surface_list_is_the_references_own():
    rc, err = syntax_only(["-DUSINGMKL", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "oracle", "gen"), "-I" + REF])
    assert rc == 0, err[:4000]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is only present in the build container")
def test_shipped_case_file_builds_unmodified_against_the_host_classes(tmp_path):
    """the reference's own veritas.cpp (its case file + main, veritas.cpp:1-170), copied next to nothing of the reference, compiles
    and links against veritas_b200/host + libveritas_b200.so; without a CUDA device the binary stops at vrt_create with the
    library's message (no CPU fallback) instead of computing anything"""
    import shutil
    lib_dir = os.path.join(ROOT, "veritas_b200")
    assert os.path.exists(os.path.join(lib_dir, "libveritas_b200.so")), "run __graft_entry__.build()"
    case = tmp_path / "case.cpp"
    shutil.copy(os.path.join(REF, "veritas.cpp"), case)          # a scratch copy: quoted includes must not find the reference's headers
    exe = tmp_path / "veritas_dropin"
    r = subprocess.run(["g++", "-O1", "-fopenmp", "-std=c++14", "-ffp-contract=off", "-w", "-I" + os.path.join(lib_dir, "host"), str(case),
                        os.path.join(lib_dir, "host", "veritas_host.cpp"), "-o", str(exe), "-L" + lib_dir, "-lveritas_b200",
                        "-Wl,--disable-new-dtags,-rpath," + lib_dir], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[:4000]
    import torch
    if torch.cuda.is_available():
        return                                                   # the full shipped run (8.5 laser periods, 5 levels) is not a unit test
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    assert run.returncode != 0 and "no CPU fallback" in (run.stdout + run.stderr)
