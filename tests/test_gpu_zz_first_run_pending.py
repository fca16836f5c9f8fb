"""First-run-pending GPU tests.  (1) Per-patch moments at the ABI (vrt_patch_moments = Rectangle::chargeR / currentR after CalculateRhoAndJ, what
Level::CollectRhoAndJ sums, Level.cpp:42-62): the level sums of the patches must reproduce the species' charge and the total
current that vrt_moments assembles, on a 3-level hierarchy from the reference's regrid (split path) and on the single-patch fused
path, and the assembled state must be left as vrt_moments leaves it.

Written after round 1's GPU budget was spent, so its first run on a B200 is the driver's: it is marked xfail(strict=False) until
it has been seen green once, and sorts last so that it cannot mask another test.  (2) The shipped five-level case through the host
classes against the reference (same status)."""
import numpy as np
import pytest

from common import load_golden, meta, species_from
import veritas_b200 as vb
from veritas_b200 import solver as S
from oracle.port import hierarchy_from_dump

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first GPU run pending (added after the round's GPU budget was spent)")]


def level_sums(ctx, s, n_finest):
    """Level::CollectRhoAndJ + Level::InterpolateRhoAndJToFinestMesh over all patches of species s"""
    charge, current = np.zeros(n_finest), np.zeros(n_finest)
    for k, p in enumerate(ctx.patches[s]):
        rtb = 2 ** p.get("depth", 0)
        cr, jr = ctx.patch_moments(s, k)
        assert cr.shape == (p["n_x"] * rtb,)
        charge[p["x_pos"] * rtb: p["x_pos"] * rtb + cr.size] += cr
        current[p["x_pos"] * rtb: p["x_pos"] * rtb + jr.size] += jr
    return charge, current


def check(ctx, n_finest):
    ctx.moments()
    before = {w: ctx.get_1d(w).copy() for w in (S.CHARGE, S.J, S.CHARGES0, S.CHARGES0 + 1)}
    total_j = np.zeros(n_finest)
    for s in range(2):
        charge, current = level_sums(ctx, s, n_finest)
        ref = before[S.CHARGES0 + s]
        assert np.abs(ref).max() > 0
        assert np.abs(charge - ref).max() <= 1e-13 * np.abs(ref).max(), (s, np.abs(charge - ref).max())
        total_j += current
    assert np.abs(total_j - before[S.J]).max() <= 1e-13 * max(np.abs(before[S.J]).max(), 1e-300)
    for w, v in before.items():                      # the assembled state (all species) is back
        assert np.array_equal(ctx.get_1d(w), v), w


def test_patch_moments_sum_to_the_species_moments_amr():
    d = load_golden("amr3_48x32_regrid")
    mt = meta(d)
    maxd = mt["Lfinest"] - 1
    sp = species_from(d)
    ctx = vb.Context(2)
    ctx.set_grid(mt["nx"] * 2 ** maxd, mt["dx"], 2, 2, 2, maxd)
    for s in range(2):
        ctx.set_species(s, sp[s]["m"], sp[s]["q"], sp[s]["pmin"], sp[s]["dp"])
    ctx.set_path(S.PATH_SPLIT)
    H = hierarchy_from_dump(d, "step0")
    for s in range(2):
        ctx.set_hierarchy(s, H[s])
    ctx.load_reference_state(d, "step0", [[p["key"] for p in h] for h in H])
    check(ctx, mt["nx"] * 2 ** maxd)
    ctx.close()


def test_patch_moments_fused_path():
    d = load_golden("single_128x64_steps")
    mt = meta(d)
    sp = species_from(d)
    ctx = vb.Context(2)
    ctx.set_grid(mt["nx"], mt["dx"], 2, 2, 2, 0)
    for s in range(2):
        ctx.set_species(s, sp[s]["m"], sp[s]["q"], sp[s]["pmin"], sp[s]["dp"])
    for s in range(2):
        ctx.set_hierarchy(s, [dict(depth=0, x_pos=0, p_pos=0, n_x=mt["nx"], n_p=mt["np"][s], up=1, down=1, left=1, right=1)])
    assert ctx.get_path(0) == S.PATH_FUSED
    ctx.load_reference_state(d, "step0")
    check(ctx, mt["nx"])
    ctx.close()


def test_shipped_five_level_case_through_host_classes(tmp_path):
    """The reference's shipped configuration (veritas.cpp:7-35: coarse 76 x 150 / 76 x 50, five levels, overdense n = 2 N_c) through
    both builds of the harness, free-running over two regrids: identical hierarchies, f and fields within tolerance.  The other
    AMR parity cases stop at three levels."""
    from test_gpu_host_layer import run_both, compare
    ref, host = run_both(tmp_path, ["76", "150", "5", "2.0", "6", "np_ion=50", "regrid_every=3", "threads=4"])
    worst, most = compare(ref, host, 6, 1e-9, 1e-12)
    print("shipped 5-level case, 6 free-running steps, regrid every 3: worst relative L2", {k: "%.2e" % v for k, v in worst.items()})
