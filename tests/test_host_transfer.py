"""Host logic on the CPU: the host-side regrid data path of the veritas_b200 host classes (the path VRT_HOST_REGRID=1 selects, and
the reference's only one) — Rectangle::GetInterpolantsREF, GetWenoValueFromCoarseLevel, GetDataFromCoarseLevelRectangle,
GetDataFromSameLevelRectangle, GetDataFromCoarseNewLevelRectangle (Rectangle.cpp:121-137, 343-415, 892-941, 1100-1128) and
Rectangle::getError with ErrorEstimate (866-890, Rectangle.hpp:128-130), Rectangle::InitializeDistribution (616-669) — on hand-made patches, bit for bit against the unmodified
reference (tests/golden/host_transfer.txt, written by oracle/_ref/ref_harness in `transfer` mode).  No Mesh, no device."""
import os
import subprocess
import pytest
from common import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
EXPECTED = os.path.join(GOLDEN, "host_transfer.txt")


def run_transfer(exe, tmp_path):
    out = tmp_path / (exe + ".txt")
    subprocess.run([os.path.join(REF_DIR, exe), "transfer", str(out)], check=True, stdout=subprocess.DEVNULL, timeout=120)
    return out.read_text().splitlines()


def test_host_transfer_and_error_flags_equal_reference_golden(tmp_path):
    assert os.path.exists(os.path.join(REF_DIR, "host_harness")), "oracle/_ref/host_harness missing: run __graft_entry__.build()"
    got, want = run_transfer("host_harness", tmp_path), open(EXPECTED).read().splitlines()
    assert len(got) == len(want) > 2500
    bad = [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, bad[:5]


def test_transfer_fixture_is_not_trivial():
    rows = open(EXPECTED).read().splitlines()
    flags = {r.split()[0]: int(r.split()[1]) for r in rows if r.startswith("flags_")}
    assert set(flags) == {"flags_coarse", "flags_new_a", "flags_old_fine"}
    assert 0 < flags["flags_coarse"] < 32 * 16 and 0 < flags["flags_new_a"] < 24 * 12 and 0 < flags["flags_old_fine"] < 16 * 8
    # the same-level copy changed the overlap of new_a (old_fine carries a shifted pattern), the coarse-new pass filled new_b
    sect, cur = {}, None
    for r in rows:
        v = r.split()
        if v[0][0].isalpha():
            cur = v[0] if len(v) == 5 and v[0] not in flags and v[0] != "interpolants" else None
            if cur:
                sect[cur] = []
        elif cur:
            sect[cur].append([float(x) for x in v[2:]])
    assert sect["after_coarse_a"] != sect["after_same_a"]
    interior_b = [row for row in sect["coarse_new_b"] if row[0] != 0.0]
    assert len(interior_b) >= 8 * 8
    # the initial distribution is a slab in x: some columns of the fine patch are empty, some are not
    nz = [row for row in sect["init_fine"] if row[0] != 0.0]
    assert 0 < len(nz) < 16 * 8 and any(row[0] != 0.0 for row in sect["init_coarse"])


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "ref_harness")), reason="reference harness not built")
def test_transfer_golden_is_what_the_reference_produces(tmp_path):
    assert run_transfer("ref_harness", tmp_path) == open(EXPECTED).read().splitlines()
