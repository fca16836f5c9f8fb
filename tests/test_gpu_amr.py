"""GPU parity tests of the AMR rows: multi-level, multi-patch hierarchies through the C ABI against the
reference's golden states and, phase by phase, against the oracle."""
import numpy as np
import pytest

from common import load_golden, rel_l2, species_from, meta
import veritas_b200 as vb
from veritas_b200 import solver as S
from oracle.port import MeshOracle, hierarchy_from_dump

pytestmark = pytest.mark.gpu
TOL = 1e-12


def laser_fn(mt):
    L = vb.load()
    return lambda t: (L.vrt_case_laser_by(mt["lam"], mt["amp"], 0.0, t), L.vrt_case_laser_bz(mt["lam"], mt["amp"], 0.0, t))


def new_ctx(d, mt):
    maxd = mt["Lfinest"] - 1
    sp = species_from(d)
    ctx = vb.Context(2)
    ctx.set_grid(mt["nx"] * 2 ** maxd, mt["dx"], 2, 2, 2, maxd)
    for s in range(2):
        ctx.set_species(s, sp[s]["m"], sp[s]["q"], sp[s]["pmin"], sp[s]["dp"])
    ctx.set_path(S.PATH_SPLIT)
    return ctx


def set_hierarchy(ctx, H):
    for s in range(2):
        ctx.set_hierarchy(s, H[s])
    return [[p["key"] for p in h] for h in H]


def strip(H):
    return [[{k: v for k, v in p.items()} for p in h] for h in H]


def test_amr_stage_parity_with_injected_phi():
    """2-level mesh, every RK stage against the reference's stage dumps (PHI injected): fine and coarse patches incl. ghosts."""
    d = load_golden("amr2_64x32_stages")
    mt = meta(d)
    ctx = new_ctx(d, mt)
    H = hierarchy_from_dump(d, "step0")
    keys = set_hierarchy(ctx, H)
    ctx.load_reference_state(d, "step0", keys)
    laser = laser_fn(mt)
    L = vb.load()
    t = float(d["step0/time"][0])
    dt = float(d["step1/dt"][0])
    worst = {}
    for i in range(6):
        tag = f"step1_stage{i}"
        ctx.moments()
        ctx.set_1d(S.PHI, d[tag + "/PHI"])
        ctx.set_scalar(S.EX0, float(d[tag + "/Ex0"][0]))
        for s in range(2):
            ctx.vlasov_stage(s, dt, i)
        t = L.vrt_update_time(t, i, dt)
        ctx.field_stage(i, dt, *laser(t))
        for s in range(2):
            for p, k in enumerate(keys[s]):
                e = rel_l2(ctx.download_f(s, p, 1), d[f"{tag}/{k}/f1"])
                worst[k] = max(worst.get(k, 0), e)
                assert e < TOL, (tag, k, e)
        for which, k in ((S.J, "J"), (S.CHARGE, "charge"), (S.A_SQUARED, "a_squared")):
            e = rel_l2(ctx.get_1d(which), d[tag + "/" + k])
            worst[k] = max(worst.get(k, 0), e)
            assert e < 1e-11, (tag, k, e)
    print("worst relative L2:", {k: "%.2e" % v for k, v in worst.items()})
    ctx.close()


@pytest.mark.parametrize("name", ["amr3_48x32_regrid", "amr2_tail_64x48_steps"])
def test_amr_per_step_parity_from_reference_state(name):
    """Protocol P1 on the hierarchies the reference's regrid produced (3 levels, up to 6 adjacent finest patches): one
    full step (own moments, own Poisson, graph replay) from each reference state; vrt_set_hierarchy is called again on the
    same context whenever the reference regridded."""
    d = load_golden(name)
    mt = meta(d)
    ctx = new_ctx(d, mt)
    laser = laser_fn(mt)
    L = vb.load()
    current, keys, compared, worst = None, None, 0, {}
    for n in range(1, mt["steps"] + 1):
        H = hierarchy_from_dump(d, f"step{n - 1}")
        if strip(H) != strip(hierarchy_from_dump(d, f"step{n}")):
            continue
        if current != strip(H):
            keys = set_hierarchy(ctx, H)
            current = strip(H)
        ctx.load_reference_state(d, f"step{n - 1}", keys)
        dt = float(d[f"step{n}/dt"][0])
        t = float(d[f"step{n - 1}/time"][0])
        lasers = []
        for i in range(6):
            t = L.vrt_update_time(t, i, dt)
            lasers += list(laser(t))
        ctx.step(dt, lasers)
        for s in range(2):
            for p, k in enumerate(keys[s]):
                f1 = ctx.download_f(s, p, 1)
                e = rel_l2(f1, d[f"step{n}/{k}/f1"])
                worst["f"] = max(worst.get("f", 0), e)
                assert e < TOL, (n, k, e)
                assert np.array_equal(ctx.download_f(s, p, 0), f1)
        for w, k in enumerate(S.FIELD_NAMES):
            e = rel_l2(ctx.download_field(w, 0), d[f"step{n}/{k}"][0])
            worst[k] = max(worst.get(k, 0), e)
            assert e < TOL, (n, k, e)
        e = rel_l2(ctx.get_1d(S.PHI), d[f"step{n}/PHI"])
        worst["PHI"] = max(worst.get("PHI", 0), e)
        assert e < 5e-8          # end to end: quasi-neutral cancellation; measured <= 3.5e-9 on these fixtures
        ctx.set_1d(S.CHARGE, d[f"step{n}/charge"])      # the solver alone, on the reference's assembled charge: <= 1e-9
        ctx.poisson()
        e = rel_l2(ctx.get_1d(S.PHI), d[f"step{n}/PHI"])
        worst["PHI_solver"] = max(worst.get("PHI_solver", 0), e)
        assert e < 1e-9, (n, "PHI from the reference's charge", e)
        compared += 1
    assert compared >= 3
    print(name, "steps compared:", compared, "worst relative L2:", {k: "%.2e" % v for k, v in worst.items()})
    ctx.close()


def test_amr_phases_against_oracle():
    """Mesh::Advance phase by phase on a 3-level hierarchy with same-level neighbours: every work plane after each sub-step
    / sync, CUDA vs the oracle, with the oracle's PHI injected so that the comparison isolates the Vlasov kernels."""
    d = load_golden("amr3_48x32_regrid")
    mt = meta(d)
    maxd = mt["Lfinest"] - 1
    n0 = max(range(mt["steps"]), key=lambda n: sum(len(h) for h in hierarchy_from_dump(d, f"step{n}")))
    H = hierarchy_from_dump(d, f"step{n0}")
    sp = species_from(d)
    O = MeshOracle(mt["nx"] * 2 ** maxd, mt["dx"], sp, H, r=2, max_depth=maxd, laser=laser_fn(mt), poisson=True)
    O.load_reference_state(d, f"step{n0}")
    ctx = new_ctx(d, mt)
    keys = set_hierarchy(ctx, H)
    ctx.load_reference_state(d, f"step{n0}", keys)
    dt = float(d[f"step{n0 + 1}/dt"][0]) if f"step{n0 + 1}/dt" in d else float(d["step1/dt"][0])
    nl = maxd + 1
    report = {}

    def compare(label, planes, state=None):
        for s in range(2):
            for p, k in enumerate(keys[s]):
                P = O.patches[s][p]
                for name in planes:
                    slot = report.get("_step", 0)
                    ref = P.a[name][slot] if name in ("FxH", "FpH") else P.a[name]
                    got = ctx.download_plane(s, p, S.PLANES.index(name), slot)
                    e = rel_l2(got, ref)
                    report[f"{label}:{name}"] = max(report.get(f"{label}:{name}", 0), e)
                    assert e < 1e-13, (label, name, k, e, np.argwhere(np.abs(got - ref) > 1e-13 * np.abs(ref).max())[:5])
                if state is not None:
                    ref = P.a["f%d" % state]
                    got = ctx.download_f(s, p, state)
                    e = rel_l2(got, ref)
                    report[f"{label}:f{state}"] = max(report.get(f"{label}:f{state}", 0), e)
                    assert e < 1e-13, (label, "f%d" % state, k, e, np.argwhere(np.abs(got - ref) > 1e-13 * np.abs(ref).max())[:5])

    for step in range(2):
        report["_step"] = step
        O.assemble(); O.update_potential()
        ctx.moments()
        for s in range(2):
            assert rel_l2(ctx.get_1d(S.CHARGES0 + s), O.charges[s]) < 1e-12
        assert rel_l2(ctx.get_1d(S.J), O.fields.J) < 1e-11
        ctx.set_scalar(S.EX0, O.fields.Ex0)       # E table = oracle's PHI and the already updated Ex0
        ctx.set_1d(S.PHI, O.fields.PHI)
        for s in range(2):
            for l in range(nl - 1, -1, -1):
                O.substep(s, l, dt, step, 0); ctx.vlasov_substep(s, l, dt, step, 0)
        compare(f"s{step}.sub0", ["ex", "ep", "fx", "fp", "FxH", "FpH", "FxDS", "FpDS"])
        for s in range(2):
            O.push_data(s, 2); ctx.push_data(s, 2)
        compare(f"s{step}.push2", [], state=2)
        for s in range(2):
            for l in range(nl - 1, -1, -1):
                O.substep(s, l, dt, step, 1); ctx.vlasov_substep(s, l, dt, step, 1)
        compare(f"s{step}.sub1", ["Rp", "Rm", "Cx", "Cp"])
        for s in range(2):
            O.push_boundary_c(s); ctx.push_boundary_c(s)
        compare(f"s{step}.pushC", ["Cx", "Cp"])
        for s in range(2):
            for l in range(nl - 1, -1, -1):
                O.substep(s, l, dt, step, 2); ctx.vlasov_substep(s, l, dt, step, 2)
            O.push_data(s, 1); ctx.push_data(s, 1)
        compare(f"s{step}.push1", [], state=1)
        O.time = O.L.vo_update_time(O.time, step, dt)
        by0, bz0 = O.laser(O.time)
        import ctypes as C
        O.L.vo_field_stage(C.byref(O.fields.c), step, dt, by0, bz0)
        ctx.field_stage(step, dt, by0, bz0)
    report.pop("_step")
    print("phase errors:", {k: "%.1e" % v for k, v in report.items()})
    ctx.close()
