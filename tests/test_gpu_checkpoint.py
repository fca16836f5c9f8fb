"""Checkpoint / restart through the C ABI: a run restarted from vrt_checkpoint_write's file must
continue bit for bit — size-independent property, no oracle involved (the reference has no checkpoint to compare with)."""
import numpy as np
import pytest

from common import load_golden, meta
import veritas_b200 as vb
from veritas_b200 import solver as S
from oracle.port import hierarchy_from_dump
from test_gpu_amr import new_ctx, set_hierarchy, laser_fn

pytestmark = pytest.mark.gpu


def snapshot(ctx, n_patches):
    out = [ctx.download_f(s, p, 1) for s in range(2) for p in range(n_patches[s])]
    out += [ctx.download_field(w, 0) for w in range(6)]
    out += [ctx.get_1d(S.PHI), ctx.get_1d(S.A_SQUARED), ctx.get_1d(S.EFIELD), np.array([ctx.get_scalar(S.EX0), ctx.get_scalar(S.TIME)])]
    return out


def test_restart_fused_bitwise(tmp_path):
    path = tmp_path / "fused.ckpt"
    kw = dict(density=0.3)
    run = vb.LaserPlasmaRun(256, 128, **kw)
    run.init_device()
    run.run_fields_phase()
    for _ in range(2):
        run.advance(run.calculate_dt())
    run.ctx.checkpoint_write(path)
    for _ in range(3):
        run.advance(run.calculate_dt())
    a = snapshot(run.ctx, [1, 1])
    run.ctx.close()

    run2 = vb.LaserPlasmaRun(256, 128, **kw)          # same Settings; no initial condition, no fields phase
    run2.ctx.checkpoint_read(path)
    run2.time = run2.ctx.get_scalar(S.TIME)
    assert run2.ctx.get_path(0) == S.PATH_FUSED
    for _ in range(3):
        run2.advance(run2.calculate_dt())
    b = snapshot(run2.ctx, [1, 1])
    run2.ctx.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_restart_amr_bitwise(tmp_path):
    """3-level hierarchy from the reference's regrid (split path): write after one step, continue two more; restart into a
    context that has never seen the hierarchy."""
    path = tmp_path / "amr.ckpt"
    d = load_golden("amr3_48x32_regrid")
    mt = meta(d)
    laser = laser_fn(mt)
    L = vb.load()
    H = hierarchy_from_dump(d, "step2")
    npatch = [len(h) for h in H]

    def steps(ctx, n, t):
        for _ in range(n):
            dt = min(0.5 * ctx.cfl_bound(), float(d["step1/dt"][0]))
            lasers = []
            for i in range(6):
                t = L.vrt_update_time(t, i, dt)
                lasers += list(laser(t))
            ctx.step(dt, lasers)
        return t

    ctx = new_ctx(d, mt)
    keys = set_hierarchy(ctx, H)
    ctx.load_reference_state(d, "step2", keys)
    for s in range(2):
        ctx.push_data(s, 1)
        ctx.commit_state(s)
    t = steps(ctx, 1, float(d["step2/time"][0]))
    ctx.checkpoint_write(path)
    steps(ctx, 2, t)
    a = snapshot(ctx, npatch)
    ctx.close()

    ctx2 = new_ctx(d, mt)
    ctx2.checkpoint_read(path, H)
    assert ctx2.get_scalar(S.TIME) == t
    steps(ctx2, 2, t)
    b = snapshot(ctx2, npatch)
    ctx2.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_checkpoint_rejects_other_grid(tmp_path):
    path = tmp_path / "g.ckpt"
    run = vb.LaserPlasmaRun(128, 64, density=0.3)
    run.init_device()
    run.ctx.checkpoint_write(path)
    run.ctx.close()
    other = vb.LaserPlasmaRun(256, 64, density=0.3)
    with pytest.raises(vb.VrtError):
        other.ctx.checkpoint_read(path)
    other.ctx.close()
