"""TEST INFRASTRUCTURE: ctypes front-end of oracle/_ref/libveritas_oracle.so (veritas_oracle.c).

`SingleLevelOracle` restates SolverManager::Advance (/root/reference/SolverManager.cpp:28-39) for
single-level meshes (one full-domain patch per species) on top of the C port.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
dbl_p = C.POINTER(C.c_double)


class VoPatch(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("n_x", "n_p", "x_pos", "p_pos", "up", "down", "left", "right", "rtb", "pad_")]
                + [(n, C.c_double) for n in ("dx", "dp", "pmin", "m", "q")]
                + [(n, dbl_p) for n in ("f0", "f1", "f2", "fx", "fp", "ex", "ep", "FxH", "FpH", "FxL", "FpL",
                                        "FxLS", "FpLS", "FxDS", "FpDS", "Rp", "Rm", "Cx", "Cp")])


class VoFields(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("x_size", "n_prepad", "n_postpad", "pad_")] + [("dx", C.c_double)]
                + [(n, dbl_p) for n in ("By", "Bz", "Ey", "Ez", "Ay", "Az", "a_squared", "PHI", "charge", "J", "neutral")]
                + [("Ex0", C.c_double)])


def build(force=False):
    so = os.path.join(_HERE, "_ref", "libveritas_oracle.so")
    src = os.path.join(_HERE, "veritas_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.vo_em_efield.restype = C.c_double
        L.vo_em_efield.argtypes = [C.POINTER(VoFields), C.c_int]
        L.vo_weno.restype = C.c_double
        L.vo_weno.argtypes = [C.c_double] * 4 + [C.c_int]
        L.vo_fct_substep.argtypes = [C.POINTER(VoPatch), C.POINTER(VoFields), C.c_double, C.c_int, C.c_int]
        L.vo_fill_domain_ghosts.argtypes = [C.POINTER(VoPatch), C.c_int]
        L.vo_mesh_advance_single.argtypes = [C.POINTER(VoPatch), C.POINTER(VoFields), C.c_double, C.c_int]
        L.vo_patch_moments.argtypes = [C.POINTER(VoPatch), C.POINTER(VoFields), C.c_void_p, dbl_p, dbl_p]
        L.vo_poisson_create.restype = C.c_void_p
        L.vo_poisson_create.argtypes = [C.c_int]
        L.vo_poisson_destroy.argtypes = [C.c_void_p]
        L.vo_update_potential.argtypes = [C.c_void_p, C.POINTER(VoFields)]
        L.vo_update_ex0.argtypes = [C.POINTER(VoFields)]
        L.vo_field_stage.argtypes = [C.POINTER(VoFields), C.c_int, C.c_double, C.c_double, C.c_double]
        L.vo_cfl_bound.restype = C.c_double
        L.vo_cfl_bound.argtypes = [C.POINTER(VoFields), C.c_int, dbl_p, dbl_p, dbl_p]
        L.vo_update_time.restype = C.c_double
        L.vo_update_time.argtypes = [C.c_double, C.c_int, C.c_double]
        pp = C.POINTER(C.POINTER(VoPatch))
        L.vo_mesh_create.restype = C.c_void_p
        L.vo_mesh_create.argtypes = [C.c_int, C.c_int, C.c_int, pp, C.POINTER(C.c_int)]
        L.vo_mesh_destroy.argtypes = [C.c_void_p]
        L.vo_mesh_get_strips.restype = C.c_int
        L.vo_mesh_get_strips.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_ubyte)]
        L.vo_mesh_get_flags.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ubyte)]
        L.vo_mesh_push_data.argtypes = [C.c_void_p, C.c_int]
        L.vo_mesh_push_boundary_c.argtypes = [C.c_void_p]
        L.vo_mesh_substep.argtypes = [C.c_void_p, C.POINTER(VoFields), C.c_int, C.c_double, C.c_int, C.c_int]
        L.vo_mesh_advance.argtypes = [C.c_void_p, C.POINTER(VoFields), C.c_double, C.c_int]
        L.vo_mesh_moments.argtypes = [C.c_void_p, C.POINTER(VoFields), dbl_p, dbl_p, dbl_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(dbl_p)


class Patch:
    """One Rectangle (Rectangle.hpp:6-22) in SoA planes."""
    PLANES = ("f0", "f1", "f2", "fx", "fp", "ex", "ep", "FxL", "FpL", "FxLS", "FpLS", "FxDS", "FpDS", "Rp", "Rm", "Cx", "Cp")

    def __init__(self, n_x, n_p, dx, dp, pmin, m, q, x_pos=0, p_pos=0, up=1, down=1, left=1, right=1, rtb=1):
        self.n_x, self.n_p = n_x, n_p
        shape = (n_x + 4, n_p + 4)
        self.a = {k: np.zeros(shape) for k in self.PLANES}
        self.a["FxH"] = np.zeros((6,) + shape)
        self.a["FpH"] = np.zeros((6,) + shape)
        self.c = VoPatch(n_x=n_x, n_p=n_p, x_pos=x_pos, p_pos=p_pos, up=up, down=down, left=left, right=right,
                         rtb=rtb, dx=dx, dp=dp, pmin=pmin, m=m, q=q)
        for k, v in self.a.items():
            setattr(self.c, k, _p(v))

    def __getattr__(self, k):
        a = self.__dict__.get("a")
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)


class Fields:
    """EMFieldSolver state (EMSolver.hpp:10-23)."""

    def __init__(self, x_size, dx, n_prepad=2, n_postpad=2):
        self.N, self.M = x_size, x_size + n_prepad + n_postpad
        self.a = {k: np.zeros((8, self.M)) for k in ("By", "Bz", "Ey", "Ez", "Ay", "Az")}
        self.a["a_squared"] = np.zeros(x_size + 1)
        for k in ("PHI", "charge", "J", "neutral"):
            self.a[k] = np.zeros(x_size)
        self.c = VoFields(x_size=x_size, n_prepad=n_prepad, n_postpad=n_postpad, dx=dx, Ex0=0.0)
        for k, v in self.a.items():
            setattr(self.c, k, _p(v))

    def __getattr__(self, k):
        a = self.__dict__.get("a")
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)

    @property
    def Ex0(self):
        return self.c.Ex0

    @Ex0.setter
    def Ex0(self, v):
        self.c.Ex0 = v

    def efield(self):
        L = lib()
        return np.array([L.vo_em_efield(C.byref(self.c), i) for i in range(self.N)])


class SingleLevelOracle:
    """SolverManager for Lfinest=1 runs: one full-domain Rectangle per species."""

    def __init__(self, n_x, n_p, dx, species, laser=None, poisson=True):
        # species: list of dicts m,q,pmin,dp,(n_p) ; laser: callable time -> (by0, bz0)
        self.L = lib()
        self.patches = [Patch(n_x, sp.get("n_p", n_p), dx, sp["dp"], sp["pmin"], sp["m"], sp["q"]) for sp in species]
        self.species = species
        self.fields = Fields(n_x, dx)
        self.time = 0.0
        self.laser = laser or (lambda t: (0.0, 0.0))
        self.charges = [np.zeros(n_x) for _ in species]
        self._scratch = np.zeros(n_x)
        self.poisson = self.L.vo_poisson_create(n_x) if poisson else None

    def __del__(self):
        if getattr(self, "poisson", None):
            self.L.vo_poisson_destroy(self.poisson)
            self.poisson = None

    def assemble(self):
        """EMFieldSolver::AssembleRhoAndJ (EMSolver.cpp:104-122)."""
        F = self.fields
        F.charge[:] = 0.0
        F.J[:] = 0.0
        for s, P in enumerate(self.patches):
            self.L.vo_patch_moments(C.byref(P.c), C.byref(F.c), None, _p(self.charges[s]), _p(self._scratch))
            self.charges[s][:] = 0.0 + (0.0 + self.charges[s])
            F.J[:] = F.J + (0.0 + self._scratch)
        for s in range(len(self.patches)):
            F.charge[:] = F.charge + self.charges[s]

    def enforce_neutralization(self):
        """EMFieldSolver::EnforceChargeNeutralization (EMSolver.cpp:621-629)."""
        self.assemble()
        self.fields.J[:] = 0.0
        self.fields.neutral[:] = -self.fields.charge

    def update_potential(self, phi_inject=None):
        F = self.fields
        if phi_inject is not None:
            F.PHI[:] = phi_inject
            self.L.vo_update_ex0(C.byref(F.c))
        else:
            self.L.vo_update_potential(self.poisson, C.byref(F.c))

    def stage(self, dt, i, phi_inject=None):
        self.assemble()
        self.update_potential(phi_inject)
        for P in self.patches:
            self.L.vo_mesh_advance_single(C.byref(P.c), C.byref(self.fields.c), dt, i)
        self.time = self.L.vo_update_time(self.time, i, dt)
        by0, bz0 = self.laser(self.time)
        self.L.vo_field_stage(C.byref(self.fields.c), i, dt, by0, bz0)

    def advance(self, dt, phi_inject=None):
        """SolverManager::Advance (SolverManager.cpp:28-39). phi_inject: optional list of 6 PHI arrays."""
        for i in range(6):
            self.stage(dt, i, None if phi_inject is None else phi_inject[i])

    def advance_fields(self, dt):
        """SolverManager::AdvanceFields (SolverManager.cpp:41-46)."""
        for i in range(6):
            self.time = self.L.vo_update_time(self.time, i, dt)
            by0, bz0 = self.laser(self.time)
            self.L.vo_field_stage(C.byref(self.fields.c), i, dt, by0, bz0)

    def cfl_bound(self):
        n = len(self.species)
        m = np.array([sp["m"] for sp in self.species]); q = np.array([sp["q"] for sp in self.species])
        dp = np.array([sp["dp"] for sp in self.species])
        return self.L.vo_cfl_bound(C.byref(self.fields.c), n, _p(m), _p(q), _p(dp))

    def load_reference_state(self, dump, tag):
        """Restart state from a ref_harness dump record group (protocol P1)."""
        F = self.fields
        for k in ("By", "Bz", "Ey", "Ez", "Ay", "Az"):
            F.a[k][:] = dump[f"{tag}/{k}"]
        F.a_squared[:] = dump[f"{tag}/a_squared"]
        F.PHI[:] = dump[f"{tag}/PHI"]
        F.charge[:] = dump[f"{tag}/charge"]
        F.J[:] = dump[f"{tag}/J"]
        F.neutral[:] = dump[f"{tag}/neutralizationCharge"]
        F.Ex0 = float(dump[f"{tag}/Ex0"][0])
        self.time = float(dump[f"{tag}/time"][0])
        for s, P in enumerate(self.patches):
            base = f"{tag}/s{s}/l0/r0/"
            if base + "f" in dump:
                f = dump[base + "f"]
                P.f0[:] = f[:, :, 0]; P.f1[:] = f[:, :, 1]; P.f2[:] = f[:, :, 2]
            else:   # reduced golden fixture: states 0 and 1 only
                P.f0[:] = dump[base + "f0"]; P.f1[:] = dump[base + "f1"]
            for k in ("FxH", "FpH"):
                if base + k in dump:
                    P.a[k][:] = np.moveaxis(dump[base + k], 2, 0)
            if base + "FxL" in dump:
                P.FxL[:] = dump[base + "FxL"][:, :, 0]
                P.FpL[:] = dump[base + "FpL"][:, :, 0]


def hierarchy_from_dump(dump, tag, n_species=2):
    """Patch descriptors per species from a ref_harness record group, in the reference's order: levels[0] (finest,
    depth 0) first, Level::rectangles order inside a level (Mesh.cpp:814)."""
    out = []
    for s in range(n_species):
        descs, l = [], 0
        while f"{tag}/s{s}/l{l}/r0/desc" in dump or any(k.startswith(f"{tag}/s{s}/l{l + 1}/") for k in dump):
            r = 0
            while f"{tag}/s{s}/l{l}/r{r}/desc" in dump:
                d = dump[f"{tag}/s{s}/l{l}/r{r}/desc"]
                descs.append(dict(depth=int(d[4]), x_pos=int(d[2]), p_pos=int(d[3]), n_x=int(d[0]), n_p=int(d[1]),
                                  up=int(d[5]), down=int(d[6]), left=int(d[7]), right=int(d[8]), key=f"s{s}/l{l}/r{r}"))
                r += 1
            l += 1
        out.append(descs)
    return out


class MeshOracle:
    """SolverManager over multi-level meshes (one vo_mesh per species): restates SolverManager::Advance
    (SolverManager.cpp:28-39) with Mesh::Advance / PushData / PushBoundaryC and the moment assembly on the C port."""

    def __init__(self, x_size_finest, dx_finest, species, hierarchies, r=2, max_depth=None, laser=None, poisson=True):
        self.L = lib()
        self.species = species
        self.r = r
        self.N = x_size_finest
        self.max_depth = max(d["depth"] for h in hierarchies for d in h) if max_depth is None else max_depth
        self.fields = Fields(x_size_finest, dx_finest)
        self.time = 0.0
        self.laser = laser or (lambda t: (0.0, 0.0))
        self.charges = [np.zeros(self.N) for _ in species]
        self._scratch = np.zeros(4 * self.N)
        self.poisson = self.L.vo_poisson_create(self.N) if poisson else None
        self.meshes, self.patches, self.descs = [], [], hierarchies
        for sp, h in zip(species, hierarchies):
            ps = []
            for d in h:
                sc = float(r) ** d["depth"]
                ps.append(Patch(d["n_x"], d["n_p"], sc * dx_finest, sc * sp["dp"], sp["pmin"], sp["m"], sp["q"], x_pos=d["x_pos"],
                                p_pos=d["p_pos"], up=d["up"], down=d["down"], left=d["left"], right=d["right"], rtb=int(r ** d["depth"])))
            arr = (C.POINTER(VoPatch) * len(ps))(*[C.pointer(p.c) for p in ps])
            depth = (C.c_int * len(ps))(*[d["depth"] for d in h])
            self.meshes.append(self.L.vo_mesh_create(len(ps), r, self.max_depth + 1, arr, depth))
            self.patches.append(ps)

    def __del__(self):
        for m in getattr(self, "meshes", []):
            self.L.vo_mesh_destroy(m)
        self.meshes = []
        if getattr(self, "poisson", None):
            self.L.vo_poisson_destroy(self.poisson)
            self.poisson = None

    def strips(self, s, p, side):
        P = self.patches[s][p]
        n = (P.n_p // self.r + 2) if side < 2 else P.n_x // self.r
        nb = (C.c_int * n)(); same = (C.c_ubyte * n)()
        self.L.vo_mesh_get_strips(self.meshes[s], p, side, nb, same)
        return list(nb), list(same)

    def flags(self, s, p):
        P = self.patches[s][p]
        out = np.zeros((P.n_x + 4, P.n_p + 4), dtype=np.uint8)
        self.L.vo_mesh_get_flags(self.meshes[s], p, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return out

    def assemble(self):
        F = self.fields
        F.charge[:] = 0.0
        F.J[:] = 0.0
        for s, m in enumerate(self.meshes):
            self.charges[s][:] = 0.0
            self.L.vo_mesh_moments(m, C.byref(F.c), _p(self.charges[s]), _p(F.J), _p(self._scratch))
        for s in range(len(self.meshes)):
            F.charge[:] = F.charge + self.charges[s]

    def update_potential(self, phi_inject=None):
        F = self.fields
        if phi_inject is not None:
            F.PHI[:] = phi_inject
            self.L.vo_update_ex0(C.byref(F.c))
        else:
            self.L.vo_update_potential(self.poisson, C.byref(F.c))

    def stage(self, dt, i, phi_inject=None):
        self.assemble()
        self.update_potential(phi_inject)
        for m in self.meshes:
            self.L.vo_mesh_advance(m, C.byref(self.fields.c), dt, i)
        self.time = self.L.vo_update_time(self.time, i, dt)
        by0, bz0 = self.laser(self.time)
        self.L.vo_field_stage(C.byref(self.fields.c), i, dt, by0, bz0)

    def advance(self, dt, phi_inject=None):
        for i in range(6):
            self.stage(dt, i, None if phi_inject is None else phi_inject[i])

    def push_data(self, s, val):
        self.L.vo_mesh_push_data(self.meshes[s], val)

    def push_boundary_c(self, s):
        self.L.vo_mesh_push_boundary_c(self.meshes[s])

    def substep(self, s, depth, dt, step, sub):
        self.L.vo_mesh_substep(self.meshes[s], C.byref(self.fields.c), depth, dt, step, sub)

    def load_reference_state(self, dump, tag):
        F = self.fields
        for k in ("By", "Bz", "Ey", "Ez", "Ay", "Az"):
            F.a[k][:] = dump[f"{tag}/{k}"]
        F.a_squared[:] = dump[f"{tag}/a_squared"]
        F.PHI[:] = dump[f"{tag}/PHI"]
        F.charge[:] = dump[f"{tag}/charge"]
        F.J[:] = dump[f"{tag}/J"]
        F.neutral[:] = dump[f"{tag}/neutralizationCharge"]
        F.Ex0 = float(dump[f"{tag}/Ex0"][0])
        self.time = float(dump[f"{tag}/time"][0])
        for s, h in enumerate(self.descs):
            for P, d in zip(self.patches[s], h):
                base = f"{tag}/{d['key']}/"
                if base + "f" in dump:
                    f = dump[base + "f"]
                    P.f0[:] = f[:, :, 0]; P.f1[:] = f[:, :, 1]; P.f2[:] = f[:, :, 2]
                else:
                    P.f0[:] = dump[base + "f0"]; P.f1[:] = dump[base + "f1"]
