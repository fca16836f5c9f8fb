"""Generate the host-logic fixtures.  (1) tests/golden/cluster_cases.txt / cluster_expected.txt: flag sets for the regrid clustering (Mesh.cpp:298-792) and what
the UNMODIFIED reference (oracle/_ref/ref_harness, `cluster` mode) makes of them.  Run in the build container:

    python tests/golden/make_host_golden.py

(2) tests/golden/settings/settings_<nx>_<np>_<Lfinest>[_<np_ion>].txt: what the reference's Settings derives for those sizes
(`settings` mode of the harness), used by tests/test_host_settings.py.

(3) tests/golden/host_transfer.txt: Rectangle::GetInterpolantsREF, the three old -> new patch transfers and Rectangle::getError on
hand-made patches (`transfer` mode), used by tests/test_host_transfer.py.

Deterministic (numpy RandomState(2017)).  The case kinds are described in oracle/ref_harness.cpp; the level sizes come from the
harness arguments NX NP LFINEST below (coarsest 32 x 16, r = 2, three levels: 32x16, 64x32, 128x64).
"""
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
NX, NP, LFINEST = 32, 16, 3


def blob(rng, w, h, n, fill):
    """n rectangles of random size inside a w x h window, each cell flagged with probability `fill`"""
    cells = set()
    for _ in range(n):
        bw, bh = rng.randint(1, max(2, w // 2)), rng.randint(1, max(2, h // 2))
        x0, p0 = rng.randint(0, w - bw + 1), rng.randint(0, h - bh + 1)
        for x in range(x0, x0 + bw):
            for p in range(p0, p0 + bh):
                if rng.rand() < fill:
                    cells.add((x, p))
    return sorted(cells)


def cases():
    rng = np.random.RandomState(2017)
    out = []
    split = lambda eff, cells: out.append("split %.17g %d " % (eff, len(cells)) + " ".join("%d %d" % c for c in cells))
    # hand-made shapes: full box, single cell, one row, one column, two boxes with a gap in x / in p / in both, an L, a cross,
    # a diagonal (no holes, inflections only), a checkerboard (bisection), a frame with an empty middle
    full = [(x, p) for x in range(3, 11) for p in range(2, 8)]
    split(0.75, full)
    split(0.75, [(7, 5)])
    split(0.75, [(x, 4) for x in range(2, 20)])
    split(0.75, [(9, p) for p in range(1, 13)])
    split(0.75, [(x, p) for x in list(range(0, 5)) + list(range(12, 18)) for p in range(3, 9)])
    split(0.75, [(x, p) for x in range(4, 10) for p in list(range(0, 3)) + list(range(9, 14))])
    split(0.75, [(x, p) for x in range(0, 4) for p in range(0, 4)] + [(x, p) for x in range(10, 16) for p in range(8, 12)])
    split(0.75, [(x, p) for x in range(0, 12) for p in range(0, 3)] + [(x, p) for x in range(0, 3) for p in range(3, 12)])
    split(0.75, [(x, p) for x in range(0, 15) for p in range(6, 9)] + [(x, p) for x in range(6, 9) for p in range(0, 15)])
    split(0.75, [(k, k) for k in range(12)])
    split(0.75, [(x, p) for x in range(10) for p in range(10) if (x + p) % 2 == 0])
    split(0.75, [(x, p) for x in range(12) for p in range(10) if x in (0, 11) or p in (0, 9)])
    # equally strong inflections on both sides of the centre (the tie-break of Mesh.cpp:455-515)
    split(0.95, [(x, p) for x in range(16) for p in range(6) if not (x in (4, 11) and p > 1)])
    split(0.95, [(x, p) for x in range(6) for p in range(16) if not (p in (4, 11) and x > 1)])
    # random blobs at several efficiencies and densities, on all three level sizes
    for (w, h) in ((32, 16), (64, 32), (128, 64)):
        for eff in (0.5, 0.75, 0.9):
            for fill in (1.0, 0.85, 0.5):
                for n in (1, 3, 6):
                    c = blob(rng, w, h, n, fill)
                    if c:
                        split(eff, c)
    # interpRectanglesUp from level 0 and level 1
    for lvl in (0, 1):
        w, h = NX * 2 ** lvl, NP * 2 ** lvl
        for _ in range(6):
            k = rng.randint(1, 5)
            boxes = []
            for _ in range(k):
                x0, p0 = rng.randint(0, w - 1), rng.randint(0, h - 1)
                boxes.append((x0, p0, rng.randint(x0, w), rng.randint(p0, h)))
            out.append("interp %d %d " % (lvl, k) + " ".join("%d %d %d %d" % b for b in boxes))
    # mergeDownFlaggedData: level-2 patches (128 x 64 index space) -> level-0 cells, incl. patches touching the domain edges
    for box in ((0, 0, 8, 8), (120, 56, 128, 64), (40, 20, 72, 36), (0, 30, 16, 34), (100, 0, 128, 6), (63, 31, 65, 33)):
        out.append("merge 0 %d %d %d %d" % box)
    for _ in range(6):
        x0, p0 = rng.randint(0, 120), rng.randint(0, 56)
        out.append("merge 0 %d %d %d %d" % (x0, p0, rng.randint(x0 + 1, 129), rng.randint(p0 + 1, 65)))
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    cases_path, exp_path = os.path.join(OUT, "cluster_cases.txt"), os.path.join(OUT, "cluster_expected.txt")
    with open(cases_path, "w") as f:
        f.write("\n".join(cases()) + "\n")
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    subprocess.check_call([HARNESS, "cluster", cases_path, exp_path, str(NX), str(NP), str(LFINEST)], stdout=subprocess.DEVNULL, env=env)
    n = sum(1 for _ in open(exp_path))
    print(f"{n} cases -> {exp_path} ({os.path.getsize(cases_path)} + {os.path.getsize(exp_path)} bytes)")
    os.makedirs(os.path.join(OUT, "settings"), exist_ok=True)
    for sizes in (("64", "32", "1"), ("32", "16", "3"), ("48", "32", "2", "16"), ("2048", "256", "1")):
        path = os.path.join(OUT, "settings", "settings_" + "_".join(sizes) + ".txt")
        subprocess.check_call([HARNESS, "settings", path] + list(sizes), stdout=subprocess.DEVNULL, env=env)
        print(path)
    # (3) host-side regrid data path on hand-made patches (`transfer` mode of the harness), tests/test_host_transfer.py
    path = os.path.join(OUT, "host_transfer.txt")
    subprocess.check_call([HARNESS, "transfer", path], stdout=subprocess.DEVNULL, env=env)
    print(path)


if __name__ == "__main__":
    main()
