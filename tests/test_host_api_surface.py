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
