"""Host logic on the CPU: the regrid clustering of the veritas_b200 host classes (veritas_b200/host/veritas_host.cpp:
getExtrema, computeSignatures, hasHole, identifyInflection, splitRectangle, interpRectanglesUp, mergeDownFlaggedData —
the reference's Mesh.cpp:298-792) against what the unmodified reference makes of the same flag sets.

tests/golden/cluster_cases.txt / cluster_expected.txt are written by tests/golden/make_host_golden.py from
oracle/_ref/ref_harness (`cluster` mode).  The host classes run the same cases through oracle/_ref/host_harness, which in this
mode creates neither a Mesh nor a device context, so the test needs no GPU.
"""
import os
import subprocess
import pytest
from common import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CASES = os.path.join(GOLDEN, "cluster_cases.txt")
EXPECTED = os.path.join(GOLDEN, "cluster_expected.txt")
SIZES = ["32", "16", "3"]           # coarsest nx, np, levels: as in make_host_golden.py


def run_cluster(exe, tmp_path):
    out = tmp_path / (exe + ".txt")
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    subprocess.run([os.path.join(REF_DIR, exe), "cluster", CASES, str(out)] + SIZES, check=True, stdout=subprocess.DEVNULL, env=env, timeout=120)
    return out.read_text().splitlines()


def describe(i):
    return open(CASES).read().splitlines()[i][:120]


def test_host_clustering_equals_reference_golden(tmp_path):
    exe = os.path.join(REF_DIR, "host_harness")
    assert os.path.exists(exe), "oracle/_ref/host_harness missing: run __graft_entry__.build()"
    got, want = run_cluster("host_harness", tmp_path), open(EXPECTED).read().splitlines()
    assert len(got) == len(want) > 100
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"case {i} ({describe(i)} ...): host classes {g[:200]} != reference {w[:200]}"


def test_clustering_covers_every_cut_kind():
    """the fixture exercises: box accepted whole, cut at a hole, cut at an inflection / bisection (several boxes without any
    empty row or column between them), and re-indexing / footprint cases"""
    cases, want = open(CASES).read().splitlines(), open(EXPECTED).read().splitlines()
    kinds = [c.split()[0] for c in cases]
    assert kinds.count("split") > 80 and kinds.count("interp") >= 10 and kinds.count("merge") >= 10
    n_boxes = [int(w.split()[0]) for c, w in zip(cases, want) if c.startswith("split")]
    assert 1 in n_boxes and 2 in n_boxes and max(n_boxes) >= 8
    # every flagged cell lies in exactly one box of its case, and no box is empty
    for c, w in zip(cases, want):
        if not c.startswith("split"):
            continue
        v = c.split()
        cells = [(int(v[3 + 2 * k]), int(v[4 + 2 * k])) for k in range(int(v[2]))]
        b = [int(x) for x in w.split()]
        boxes = [tuple(b[1 + 4 * k: 5 + 4 * k]) for k in range(b[0])]
        hits = [0] * len(boxes)
        for (x, p) in cells:
            inside = [k for k, (x0, p0, x1, p1) in enumerate(boxes) if x0 <= x <= x1 and p0 <= p <= p1]
            assert len(inside) == 1, (c[:80], x, p, inside)
            hits[inside[0]] += 1
        assert all(hits), c[:80]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "ref_harness")), reason="reference harness not built")
def test_cluster_golden_is_what_the_reference_produces(tmp_path):
    assert run_cluster("ref_harness", tmp_path) == open(EXPECTED).read().splitlines()
