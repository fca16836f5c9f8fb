"""Host logic on the CPU: what the veritas_b200 host `Settings` (veritas_b200/host/veritas_host.cpp) derives from Input / Particles
— level sizes and spacings, species constants, fMax (DetermineMaximum), the stage times of UpdateTime — against the reference's
Settings (Settings.cpp:5-195) on the same case file, printed with 17 significant digits and compared as text.

tests/golden/settings/*.txt are outputs of oracle/_ref/ref_harness (`settings` mode, unmodified reference); the host classes run
through oracle/_ref/host_harness, which creates no device context in this mode.
"""
import glob
import os
import subprocess
import pytest
from common import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "settings", "settings_*.txt")))


def run_settings(exe, args, tmp_path):
    out = tmp_path / (exe + ".txt")
    subprocess.run([os.path.join(REF_DIR, exe), "settings", str(out)] + args, check=True, stdout=subprocess.DEVNULL, timeout=120)
    return out.read_text().splitlines()


def fixture_args(path):
    return os.path.basename(path)[len("settings_"):-len(".txt")].split("_")


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_host_settings_equal_reference_golden(path, tmp_path):
    assert os.path.exists(os.path.join(REF_DIR, "host_harness")), "oracle/_ref/host_harness missing: run __graft_entry__.build()"
    got, want = run_settings("host_harness", fixture_args(path), tmp_path), open(path).read().splitlines()
    assert len(want) > 50
    assert got == want, [(g, w) for g, w in zip(got, want) if g != w][:5]


def test_settings_fixtures_cover_levels_and_unequal_species():
    assert len(FIXTURES) >= 4
    text = {os.path.basename(p): open(p).read() for p in FIXTURES}
    assert any("GetDx2 " in t for t in text.values())                     # a three-level case
    assert any("p_size0 32" in t and "p_size1 16" in t for t in text.values())   # species with different p resolution
    for t in text.values():                                               # UpdateTime ends a step at t + dt (c = 0 .5 .332 .62 .85 1)
        v = dict(line.split() for line in t.splitlines())
        assert float(v["time_step0_stage5"]) == pytest.approx(1.25e-18, rel=1e-15)
        assert float(v["time_step1_stage5"]) == pytest.approx(1.25e-18 + 3.0e-17, rel=1e-15)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "ref_harness")), reason="reference harness not built")
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_settings_golden_is_what_the_reference_prints(path, tmp_path):
    assert run_settings("ref_harness", fixture_args(path), tmp_path) == open(path).read().splitlines()
