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


def test_restart_through_the_host_classes(tmp_path):
    """SolverManager::Checkpoint / SolverManager::Restart (additive members of the host classes; the reference's driver has no restart
    entry): the reference-style driver loop on a 3-level hierarchy with a regrid every 3 steps, once straight through 8 steps
    writing a checkpoint after step 4, once restarted from that checkpoint into a freshly constructed SolverManager (Level /
    Rectangle objects rebuilt from the device's descriptors, Mesh::AdoptDeviceHierarchy) — steps 5..8, which cross a regrid that
    starts from the adopted hierarchy, must reproduce the uninterrupted run bit for bit."""
    import os
    import subprocess
    from oracle.dumpio import read_dump
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "oracle", "_ref", "host_harness")
    if not os.path.exists(host):
        pytest.fail(f"{host} missing: run __graft_entry__.build() in the build container")
    args = ["48", "32", "3", "0.5", "8", "pre_steps=1600", "regrid_every=3", "threads=4"]
    env = dict(os.environ, VRT_HARNESS_CKPT=str(tmp_path / "run.ckpt"), OPENBLAS_NUM_THREADS="1")
    for name, extra in (("straight", ["ckpt_write=4"]), ("restarted", ["ckpt_restart=1"])):
        r = subprocess.run([host, str(tmp_path / f"{name}.bin")] + args + extra, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
    a, b = read_dump(str(tmp_path / "straight.bin")), read_dump(str(tmp_path / "restarted.bin"))
    compared = 0
    for n in range(5, 9):
        keys_a = sorted(k for k in a if k.startswith(f"step{n}/"))
        assert keys_a == sorted(k for k in b if k.startswith(f"step{n}/")) and keys_a, n       # same records: same hierarchy after the regrid
        for k in keys_a:
            assert np.array_equal(a[k], b[k]), k
            compared += 1
    assert "step4/time" not in b and compared > 100
    print(f"restart through the host classes: {compared} records of steps 5..8 bitwise equal")


def test_checkpoint_read_is_transactional(tmp_path):
    """vrt_checkpoint_read validates the whole file before it touches the context: a truncated file, trailing garbage, a corrupt
    patch count and a checkpoint of another grid are all refused with an error code, nothing throws across the C boundary, and the
    context keeps its state bit for bit (it continues exactly like an undisturbed twin); a good file still restores afterwards."""
    import struct
    good = tmp_path / "good.ckpt"
    kw = dict(density=0.3)

    def fresh():
        run = vb.LaserPlasmaRun(128, 64, **kw)
        run.init_device()
        run.time = 3 * run.T
        run.advance(run.calculate_dt())
        return run

    run, twin = fresh(), fresh()
    run.ctx.checkpoint_write(good)
    raw = open(good, "rb").read()
    header = 8 + 10 * 4 + 2 * 8          # magic, ten ints, dx, time
    fields = (6 * 8 * run.ctx.M + (run.ctx.N + 1) + 2 * run.ctx.N + 1) * 8
    n_off = header + fields + 4 * 8 + 4  # species record: VrtSpecies (4 doubles), path, then the patch count
    bad = {
        "truncated": raw[: len(raw) // 2],
        "trailing": raw + b"\0" * 24,
        "patch_count": raw[:n_off] + struct.pack("<i", 2 ** 30) + raw[n_off + 4:],
        "zero_patches": raw[:n_off] + struct.pack("<i", 0) + raw[n_off + 4:],
        "not_a_checkpoint": b"VRTCKPT0" + raw[8:],
    }
    other = vb.LaserPlasmaRun(64, 64, **kw)
    other.init_device()
    other.ctx.checkpoint_write(tmp_path / "other_grid.ckpt")
    other.ctx.close()
    for name, blob in bad.items():
        p = tmp_path / f"{name}.ckpt"
        open(p, "wb").write(blob)
        with pytest.raises(vb.VrtError):
            run.ctx.checkpoint_read(p)
    with pytest.raises(vb.VrtError):
        run.ctx.checkpoint_read(tmp_path / "other_grid.ckpt")
    # the refused reads left the context untouched: it continues like its twin
    for r in (run, twin):
        for _ in range(2):
            r.advance(r.calculate_dt())
    for x, y in zip(snapshot(run.ctx, [1, 1]), snapshot(twin.ctx, [1, 1])):
        assert np.array_equal(x, y)
    twin.ctx.close()
    run.ctx.checkpoint_read(good)
    run.time = run.ctx.get_scalar(S.TIME)
    run.advance(run.calculate_dt())
    run.ctx.close()


def test_failed_set_hierarchy_leaves_no_half_built_species():
    """A refused or failed vrt_set_hierarchy / vrt_regrid must not leave a species with descriptors but no storage: every later
    hot-path call answers with an error code (VRT_ERR_STATE) instead of touching freed planes, and a valid hierarchy can be set
    afterwards."""
    d = load_golden("amr3_48x32_regrid")
    mt = meta(d)
    ctx = new_ctx(d, mt)
    H = hierarchy_from_dump(d, "step2")
    keys = set_hierarchy(ctx, H)
    ctx.load_reference_state(d, "step2", keys)
    ctx.moments()
    # (1) a descriptor vrt_set_hierarchy refuses (odd size on a refined level): the species holds nothing afterwards
    bad = [dict(p) for p in H[0]]
    bad[0]["n_x"] += 1
    with pytest.raises(vb.VrtError):
        ctx.set_hierarchy(0, bad)
    for call in (ctx.moments, lambda: ctx.push_data(0, 1), lambda: ctx.download_f(0, 0, 1), lambda: ctx.step(1e-18, [0.0] * 12)):
        with pytest.raises(vb.VrtError):
            call()
    # (2) a valid hierarchy afterwards: the context works again
    ctx.set_hierarchy(0, H[0])
    ctx.load_reference_state(d, "step2", keys)
    ctx.moments()
    # (3) a refused vrt_regrid keeps the resident hierarchy and its data
    before = [ctx.download_f(0, k, 1) for k in range(len(H[0]))]
    with pytest.raises(vb.VrtError):
        ctx.regrid(0, bad)
    ctx.patches[0] = [{k: p[k] for k in S.DESC_KEYS} for p in H[0]]      # the harness's own copy of the descriptors
    for k, a in enumerate(before):
        assert np.array_equal(ctx.download_f(0, k, 1), a)
    ctx.moments()
    ctx.close()
