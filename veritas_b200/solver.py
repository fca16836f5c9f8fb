"""Python front-end over the C ABI: `Context` (one vrt_ctx) and `LaserPlasmaRun`, which mirrors the
reference's driver loop (veritas.cpp:118-170) and SolverManager (SolverManager.cpp:8-46, 162-164) for the
shipped laser-plasma case on a single-level mesh.  Test/bench harness only — the C++ host classes live in
veritas_b200/host/."""
import ctypes as C
import numpy as np

from ._lib import load, VrtError, PatchDesc, CaseParams, CaseDerived, dbl_p

BY, BZ, EY, EZ, AY, AZ = range(6)
PHI, CHARGE, J, A_SQUARED, NEUTRALIZATION, EFIELD = range(6)
CHARGES0 = 16
EX0, TIME = 0, 1
PATH_AUTO, PATH_SPLIT, PATH_FUSED = 0, 1, 2
FIELD_NAMES = ("By", "Bz", "Ey", "Ez", "Ay", "Az")
PLANES = ("FxH", "FpH", "FxL", "FpL", "FxDS", "FpDS", "Rp", "Rm", "Cx", "Cp", "ex", "ep", "fx", "fp")
DESC_KEYS = ("depth", "x_pos", "p_pos", "n_x", "n_p", "up", "down", "left", "right")

M_E = 9.10938291e-31      # veritas.cpp:16
Q_E = 1.60217657e-19      # veritas.cpp:17
EPS0_INV = 1.1294e+11      # veritas.hpp:22 (literal)
CS = 299792458.0          # veritas.hpp:26


def _p(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(dbl_p)


class Context:
    def __init__(self, n_species=2, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.vrt_create(C.byref(h), device, n_species)
        if rc:
            raise VrtError(f"vrt_create failed ({rc}): {self.L.vrt_global_error().decode()}")
        self.h = h
        self.n_species = n_species
        self.N = self.M = 0
        self.patches = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.vrt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise VrtError(f"{what} failed ({rc}): {self.L.vrt_last_error(self.h).decode()}")

    def call(self, name, *args):
        self._ck(getattr(self.L, name)(self.h, *args), name)

    # configuration
    def set_grid(self, N, dx, pre=2, post=2, r=2, max_depth=0):
        self.call("vrt_set_grid", N, dx, pre, post, r, max_depth)
        self.N, self.M = N, N + pre + post

    def set_species(self, s, m, q, pmin, dp):
        self.call("vrt_set_species", s, m, q, pmin, dp)

    def set_hierarchy(self, s, patches):
        patches = [{k: p[k] for k in DESC_KEYS} for p in patches]
        arr = (PatchDesc * len(patches))(*[PatchDesc(**p) for p in patches])
        self.call("vrt_set_hierarchy", s, len(patches), arr)
        self.patches[s] = patches

    def set_path(self, path):
        self.call("vrt_set_path", path)

    def patch_energy(self, s, patch):
        """Rectangle::CalculateEnergy of one patch: n_p * r^depth values"""
        p = self.patches[s][patch]
        out = np.zeros(p["n_p"] * 2 ** p.get("depth", 0))
        self.call("vrt_patch_energy", s, patch, _p(out))
        return out

    def patch_moments(self, s, patch):
        """Rectangle::chargeR, currentR of one patch after CalculateRhoAndJ: n_x * r^depth values each"""
        p = self.patches[s][patch]
        n = p["n_x"] * 2 ** p.get("depth", 0)
        charge, current = np.zeros(n), np.zeros(n)
        self.call("vrt_patch_moments", s, patch, _p(charge), _p(current))
        return charge, current

    def checkpoint_write(self, path):
        self.call("vrt_checkpoint_write", str(path).encode())

    def checkpoint_read(self, path, patches=None):
        """patches[s] = descriptors of the stored hierarchy (only needed by this harness's upload/download shape checks;
        the library recreates the hierarchy from the file)."""
        self.call("vrt_checkpoint_read", str(path).encode())
        if patches is not None:
            for s, ps in enumerate(patches):
                self.patches[s] = [{k: p[k] for k in DESC_KEYS} for p in ps]

    def regrid(self, s, patches):
        """vrt_regrid: Mesh::InterMeshDataTransfer between the resident hierarchy and `patches` on the device"""
        patches = [{k: p[k] for k in DESC_KEYS} for p in patches]
        arr = (PatchDesc * len(patches))(*[PatchDesc(**p) for p in patches])
        self.call("vrt_regrid", s, len(patches), arr)
        self.patches[s] = patches

    def fused_plan(self, s):
        out = (C.c_int * 6)()
        self.call("vrt_fused_plan", s, out)
        return dict(W=out[0], strips=out[1], chunks=out[2], interior_ctas=out[3], moments_cpt=out[4], moments_threads=out[5])

    def get_path(self, s):
        return self.L.vrt_get_path(self.h, s)

    # data
    def upload_f(self, s, patch, state, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        p = self.patches[s][patch]
        assert arr.shape == (p["n_x"] + 4, p["n_p"] + 4), arr.shape
        self.call("vrt_patch_upload_f", s, patch, state, _p(arr))

    def download_f(self, s, patch, state, out=None):
        p = self.patches[s][patch]
        if out is None:
            out = np.zeros((p["n_x"] + 4, p["n_p"] + 4))
        self.call("vrt_patch_download_f", s, patch, state, _p(out))
        return out

    def download_plane(self, s, patch, which, slot=0):
        """One work plane (PLANES index) of a patch, padded layout; split path only."""
        p = self.patches[s][patch]
        out = np.zeros((p["n_x"] + 4, p["n_p"] + 4))
        self.call("vrt_patch_download_plane", s, patch, which, slot, _p(out))
        return out

    def upload_field(self, which, slot, arr):
        self.call("vrt_field_upload", which, slot, _p(np.ascontiguousarray(arr, dtype=np.float64)))

    def download_field(self, which, slot, out=None):
        if out is None:
            out = np.zeros(self.M)
        self.call("vrt_field_download", which, slot, _p(out))
        return out

    def set_1d(self, which, arr):
        self.call("vrt_set_1d", which, _p(np.ascontiguousarray(arr, dtype=np.float64)))

    def get_1d(self, which, out=None):
        """out: optional preallocated (e.g. pinned) float64 buffer of N (A_SQUARED: N + 1) entries"""
        if out is None:
            out = np.zeros(self.N + 1 if which == A_SQUARED else self.N)
        self.call("vrt_get_1d", which, _p(out))
        return out

    def set_scalar(self, which, v):
        self.call("vrt_set_scalar", which, float(v))

    def get_scalar(self, which):
        v = C.c_double()
        self.call("vrt_get_scalar", which, C.byref(v))
        return v.value

    # hot path
    def moments(self): self.call("vrt_moments")
    def enforce_neutralization(self): self.call("vrt_enforce_neutralization")
    def poisson(self): self.call("vrt_poisson")
    def vlasov_stage(self, s, dt, step): self.call("vrt_vlasov_stage", s, float(dt), step)
    def vlasov_substep(self, s, depth, dt, step, sub): self.call("vrt_vlasov_substep", s, depth, float(dt), step, sub)
    def push_data(self, s, val): self.call("vrt_push_data", s, val)
    def push_boundary_c(self, s): self.call("vrt_push_boundary_c", s)
    def field_stage(self, step, dt, by0, bz0): self.call("vrt_field_stage", step, float(dt), float(by0), float(bz0))
    def commit_state(self, s): self.call("vrt_commit_state", s)
    def sync(self): self.call("vrt_sync")

    def cfl_bound(self):
        v = C.c_double()
        self.call("vrt_cfl_bound", C.byref(v))
        return v.value

    def step(self, dt, laser):
        a = (C.c_double * 12)(*laser)
        self.call("vrt_step", float(dt), a)

    def step_fields(self, dt, laser):
        a = (C.c_double * 12)(*laser)
        self.call("vrt_step_fields", float(dt), a)

    def last_step_launches(self):
        return self.L.vrt_last_step_launches(self.h)

    def load_reference_state(self, dump, tag, keys=None):
        """keys[s][patch] = record key of each patch (AMR hierarchies, oracle.port.hierarchy_from_dump).
        Restart from an oracle/ref_harness dump (SURVEY.md H0 protocol P1): f per patch, the six transverse
        arrays (all slots), PHI, Ex0, neutralizationCharge, charge, J, a_squared, time."""
        for w, k in enumerate(FIELD_NAMES):
            a = dump[f"{tag}/{k}"]
            for slot in range(8):
                self.upload_field(w, slot, a[slot])
        self.set_1d(A_SQUARED, dump[f"{tag}/a_squared"])
        self.set_1d(CHARGE, dump[f"{tag}/charge"])
        self.set_1d(J, dump[f"{tag}/J"])
        self.set_1d(NEUTRALIZATION, dump[f"{tag}/neutralizationCharge"])
        self.set_scalar(EX0, float(dump[f"{tag}/Ex0"][0]))
        self.set_1d(PHI, dump[f"{tag}/PHI"])
        self.set_scalar(TIME, float(dump[f"{tag}/time"][0]))
        for s in range(self.n_species):
            for r, _ in enumerate(self.patches[s]):
                base = f"{tag}/" + (keys[s][r] if keys else f"s{s}/l0/r{r}") + "/"
                if base + "f" in dump:
                    f0, f1 = dump[base + "f"][:, :, 0], dump[base + "f"][:, :, 1]
                else:
                    f0, f1 = dump[base + "f0"], dump[base + "f1"]
                self.upload_f(s, r, 0, f0)
                self.upload_f(s, r, 1, f1)


def case_params(**kw):
    d = dict(lambda_=1e-6, a0=1.0, density=2.0, temp_frac=5e-4, pmax_e=20.0, pmax_i=200.0, box_lambdas=10.0, ion_mass_ratio=1836.0)
    d.update(kw)
    return CaseParams(**d)


class LaserPlasmaRun:
    """The shipped laser-plasma case (veritas.cpp:7-115) on a single-level mesh: two species (e-, p+ x1836)."""

    def __init__(self, nx, np_e, np_i=None, density=2.0, a0=1.0, device=0, path=PATH_AUTO, cfl=0.5,
                 slab=None, graph=True, xl=3.0e-6, xr=7.0e-6):
        self.L = load()
        np_i = np_i or np_e
        self.nx, self.np = nx, (np_e, np_i)
        self.cfl = cfl
        self.m = (M_E, M_E * 1836)
        self.q = (-Q_E, Q_E)
        cp = case_params(density=density, a0=a0)
        cd = CaseDerived()
        ps = (C.c_uint * 2)(np_e, np_i)
        rc = self.L.vrt_case_derive(C.byref(cp), self.m[0], self.q[0], nx, ps, 1e-8, C.byref(cd))
        if rc:
            raise VrtError("vrt_case_derive failed")
        self.cd = cd
        self.dx = cd.dx
        self.lam, self.amp = cd.tempEM[0], cd.tempEM[1]
        self.T = self.lam / CS
        self.xl, self.xr = xl, xr
        self.ctx = Context(2, device)
        self.ctx.set_grid(nx, self.dx, 2, 2, 2, 0)     # preLength = postLength = 0 -> pads of 2 (EMSolver.cpp:7-8)
        for s in range(2):
            self.ctx.set_species(s, self.m[s], self.q[s], cd.pmin[s], cd.dp[s])
        if slab is not None:
            from .parallel import slab_bounds
            rank, n_ranks = slab
            self.ctx.call("vrt_set_slab", rank, n_ranks, *slab_bounds(nx, rank, n_ranks))
        self.ctx.set_path(path)
        self.ctx.call("vrt_set_option", 0, 1 if graph else 0)
        for s in range(2):
            self.ctx.set_hierarchy(s, [dict(depth=0, x_pos=0, p_pos=0, n_x=nx, n_p=self.np[s], up=1, down=1, left=1, right=1)])
        self.time = 0.0
        self.dt_max = self.T / 400

    def laser(self, t):
        return (self.L.vrt_case_laser_by(self.lam, self.amp, 0.0, t), self.L.vrt_case_laser_bz(self.lam, self.amp, 0.0, t))

    def stage_lasers(self, dt):
        """GetBY/GetBZ(0, time) after each UpdateTime(i) (SolverManager.cpp:37-38)."""
        out, t = [], self.time
        for i in range(6):
            t = self.L.vrt_update_time(t, i, dt)
            out += list(self.laser(t))
        return out, t

    def init_device(self):
        """Initial condition on the device + SolverManager ctor tail (PushData, EnforceChargeNeutralization)."""
        for s in range(2):
            self.ctx.call("vrt_init_maxwellian_slab", s, self.xl, self.xr, self.cd.temp0[s], self.cd.temp1[s], self.cd.quadratureDepth)
        self.ctx.enforce_neutralization()

    def advance_fields(self, dt):
        lasers, t = self.stage_lasers(dt)
        self.ctx.step_fields(dt, lasers)
        self.time = t

    def advance(self, dt):
        lasers, t = self.stage_lasers(dt)
        self.ctx.step(dt, lasers)
        self.time = t

    def calculate_dt(self):
        return min(self.cfl * self.ctx.cfl_bound(), self.dt_max)

    def run_fields_phase(self):
        """veritas.cpp:135-144: fields-only while t <= 3T.  The reference's driver uses the fixed dt = T/400 there, which is
        only stable while c dt/dx < ~1 (nx <~ 4000 for its 10-wavelength box); like its plasma phase, this driver bounds the
        step by SolverManager::CalculateDt so that fine meshes (config 3: c T/400 = 16 dx) stay finite."""
        t, n = self.dt_max, 0
        while not (t > 3 * self.T):
            dt = self.calculate_dt()
            self.advance_fields(dt)
            t += dt
            n += 1
        return n

    def cells(self):
        return self.nx * (self.np[0] + self.np[1])


def connectivity(patches, r=2, max_depth=0):
    """Host-side connectivity tables of a hierarchy (vrt_conn_*; needs no device): list per patch of
    dict(nb=[4 lists], same=[4 lists], flags=uint8 array)."""
    L = load()
    patches = [{k: p[k] for k in DESC_KEYS} for p in patches]
    arr = (PatchDesc * len(patches))(*[PatchDesc(**p) for p in patches])
    h = C.c_void_p()
    rc = L.vrt_conn_create(C.byref(h), len(patches), arr, r, max_depth)
    if rc:
        raise VrtError(f"vrt_conn_create failed ({rc})")
    out = []
    for k, p in enumerate(patches):
        nb, same = [], []
        for side in range(4):
            n = (p["n_p"] // r + 2) if side < 2 else p["n_x"] // r
            a = (C.c_int * n)(); b = (C.c_ubyte * n)()
            assert L.vrt_conn_strips(h, k, side, a, b) == n
            nb.append(list(a)); same.append(list(b))
        fl = np.zeros((p["n_x"] + 4, p["n_p"] + 4), dtype=np.uint8)
        assert L.vrt_conn_flags(h, k, fl.ctypes.data_as(C.POINTER(C.c_ubyte))) == 0
        out.append(dict(nb=nb, same=same, flags=fl))
    L.vrt_conn_destroy(h)
    return out
