"""Shared helpers for the test-suite: golden fixtures, norms, case set-up."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k.replace("|", "/"): z[k] for k in z.files}


def rel_l2(a, b):
    """relative L2 of a against the reference b (absolute if b == 0)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


def species_from(dump):
    return [dict(m=dump[f"species{s}"][0], q=dump[f"species{s}"][1], pmin=dump[f"species{s}"][2], dp=dump[f"species{s}"][3])
            for s in range(2)]


def meta(dump):
    nx, np_, lf, dens, steps, npre, dt0, t = dump["meta"]
    n_p = [dump[f"step0/s{s}/l0/r0/f0"].shape[1] - 4 for s in range(2)]
    return dict(nx=int(nx), np=n_p, Lfinest=int(lf), density=float(dens), steps=int(steps), dx=float(dump["dx"][0]),
                lam=float(dump["tempEM"][0]), amp=float(dump["tempEM"][1]))
