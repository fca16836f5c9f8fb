"""Per-patch moments at the ABI: vrt_patch_moments = Rectangle::chargeR / currentR after Rectangle::CalculateRhoAndJ
(Rectangle.cpp:157-282), what Level::CollectRhoAndJ sums (Level.cpp:42-62).  (1) Against the reference itself: the harness
(patch_moments=1) runs EMFieldSolver::AssembleRhoAndJ on the dumped state and dumps every patch's chargeR / currentR; the kernel
must reproduce them patch by patch on a 3-level hierarchy from the reference's regrid (coarse patches interpolate to rtb
sub-cells and skip nested cells) and on the single-patch fused path.  (2) Self-consistency: the level sums of the patches
reproduce the species' charge and the total current that vrt_moments assembles, and the assembled state is left as
vrt_moments leaves it."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import load_golden, meta, species_from
import veritas_b200 as vb
from veritas_b200 import solver as S
from oracle.port import hierarchy_from_dump

pytestmark = pytest.mark.gpu


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


def against_reference(tmp_path, args, path):
    from test_gpu_reference_on_box import run_reference
    d = run_reference(tmp_path, args + ["patch_moments=1"])
    mt = meta(d)
    maxd = mt["Lfinest"] - 1
    sp = species_from(d)
    steps = int(d["meta"][4])
    worst = {"chargeR": 0.0, "currentR": 0.0}
    n_patches = 0
    for tag in (f"step{n}" for n in range(steps + 1) if f"step{n}/time" in d):
        ctx = vb.Context(2)
        ctx.set_grid(mt["nx"] * 2 ** maxd, mt["dx"], 2, 2, 2, maxd)
        for s in range(2):
            ctx.set_species(s, sp[s]["m"], sp[s]["q"], sp[s]["pmin"], sp[s]["dp"])
        ctx.set_path(path)
        H = hierarchy_from_dump(d, tag)
        for s in range(2):
            ctx.set_hierarchy(s, H[s])
        ctx.load_reference_state(d, tag, [[p["key"] for p in h] for h in H])
        for s in range(2):
            for k, p in enumerate(H[s]):
                cr, jr = ctx.patch_moments(s, k)
                for name, got in (("chargeR", cr), ("currentR", jr)):
                    ref = d[f"{tag}/{p['key']}/{name}"]
                    assert got.shape == ref.shape, (tag, p["key"], name, got.shape, ref.shape)
                    scale = max(np.abs(d[f"{tag}/{q['key']}/{name}"]).max() for q in H[s])      # the species' largest entry on any patch
                    e = float(np.abs(got - ref).max() / scale) if scale > 0 else float(np.abs(got).max())
                    worst[name] = max(worst[name], e)
                    assert e < 1e-12, (tag, p["key"], name, e)
                n_patches += 1
        ctx.close()
    return worst, n_patches


def test_patch_moments_against_reference_three_levels(tmp_path):
    worst, n = against_reference(tmp_path, ["48", "32", "3", "0.5", "8", "pre_steps=1600", "regrid_every=2", "dump_every=4", "threads=4"], S.PATH_SPLIT)
    assert n >= 3 * 2 * 3
    print("vrt_patch_moments vs Rectangle::chargeR/currentR, 3 levels, %d patches: max |diff| / max|ref|" % n, {k: "%.2e" % v for k, v in worst.items()})


def test_patch_moments_against_reference_fused(tmp_path):
    worst, n = against_reference(tmp_path, ["128", "64", "1", "0.1", "2", "pre_steps=2000", "threads=4"], S.PATH_FUSED)
    print("vrt_patch_moments vs Rectangle::chargeR/currentR, fused path: max |diff| / max|ref|", {k: "%.2e" % v for k, v in worst.items()})
