"""Generate the golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_harness, built by oracle/Makefile
from /root/reference).  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Fixtures are compressed .npz files of ref_harness records.  To stay small, stage records keep only state 1 of f
(the current stage value); step records keep states 0 and 1.
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.dumpio import read_dump  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (nx, np, Lfinest, density, steps, extra args)
    # laser well inside the slab (pre_steps=2000 -> t = 5T), underdense n = 0.1 N_c
    "single_64x32_stages": (64, 32, 1, 0.1, 2, ["stage_dumps=1", "pre_steps=2000", "threads=1"]),
    # shipped overdense density, the shipped fields-only phase (t = 3T), per-step records
    "single_128x64_steps": (128, 64, 1, 2.0, 6, ["threads=1"]),
    # unequal species resolution
    "single_96x48x24_steps": (96, 48, 1, 0.5, 3, ["np_ion=24", "pre_steps=1600", "threads=1"]),
    # 2-level AMR (BASELINE config 2 shape: coarse 64x32, one fine patch), per-stage records with PHI for injection
    "amr2_64x32_stages": (64, 32, 2, 0.5, 1, ["stage_dumps=1", "pre_steps=1600", "threads=1"]),
    # 3-level AMR with regridding every 2 steps (BASELINE config 4 shape): up to 6 adjacent finest patches
    "amr3_48x32_regrid": (48, 32, 3, 0.5, 9, ["pre_steps=1600", "regrid_every=2", "threads=1"]),
    # 2-level AMR, refinement forced in the high-momentum tail (config 2: "AMR in the high-momentum tail")
    "amr2_tail_64x48_steps": (64, 48, 2, 0.3, 3, ["pre_steps=1700", "refine_mode=1", "tail_p0=1", "threads=1"]),
}


def reduce(records):
    out = {}
    for k, v in records.items():
        if k.endswith("/f"):
            if "_stage" in k:
                out[k + "1"] = v[:, :, 1].copy()
            else:
                out[k + "0"] = v[:, :, 0].copy()
                out[k + "1"] = v[:, :, 1].copy()
        else:
            out[k] = v
    return out


def main(only=None):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    for name, (nx, np_, lf, dens, steps, extra) in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "dump.bin")
            cmd = [HARNESS, path, str(nx), str(np_), str(lf), str(dens), str(steps)] + extra
            env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, env=env)
            rec = reduce(read_dump(path))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k.replace("/", "|"): v for k, v in rec.items()})
        print(name, "%.1f KB" % (os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


if __name__ == "__main__":
    main(sys.argv[1:])
