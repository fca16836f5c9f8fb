#!/usr/bin/env python
"""Count the instructions of the fused stage kernel's interior x-loop in the SASS of the built object (no GPU needed):

    python tools/sass_count.py [--obj veritas_b200/build/vrt_fused.cu.o] [--version v17] [--write]

For every instance k_fused_stage<S, U, 128> it finds the innermost-largest loop of the interior body (the x-loop, unrolled U
times), and reports per thread and column: all instructions, fp64-pipe instructions (DADD DMUL DFMA DSETP), and the opcode mix.
--write stores profiles/fused_sass_counts.json, which bench.py uses as the numerator of the fp64 roofline (roofline.fp64)."""
import argparse
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DADD", "DMUL", "DFMA", "DSETP")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0]


def loops(body):
    out = []
    for addr, text in body:
        if opcode(text) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                out.append((int(m.group(1), 16), addr))
    return out


def analyse(body, unroll):
    """the x-loops are the two largest loops (EDGE = true / false bodies); the interior one has fewer instructions (no predicates)"""
    ls = sorted(loops(body), key=lambda l: l[1] - l[0], reverse=True)
    cands = []
    for lo, hi in ls[:4]:
        ins = [t for a, t in body if lo <= a <= hi]
        ops = collections.Counter(opcode(t) for t in ins)
        if ops.get("BAR", 0) >= 2 * unroll:          # two barriers per column
            cands.append((len(ins), ops))
    n, ops = min(cands, key=lambda c: c[0])
    cols = ops["BAR"] // 2
    fp64 = sum(ops[o] for o in FP64)
    return {"columns_per_iteration": cols, "instr_per_column": n / cols, "fp64_per_column": fp64 / cols,
            "mix_per_iteration": {k: v for k, v in ops.most_common(24)},
            "local_memory_ops_per_iteration": ops.get("LDL", 0) + ops.get("STL", 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obj", default=os.path.join(ROOT, "veritas_b200", "build", "vrt_fused.cu.o"))
    ap.add_argument("--version", default=None)
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    res = {}
    for name, body in functions(args.obj):
        m = re.search(r"k_fused_stageILi(\d)ELi(\d)ELi(\d+)E(?:Lb([01])E)?", name)
        if not m or m.group(3) != "128":
            continue
        S, U, lean = int(m.group(1)), int(m.group(2)), m.group(4) == "1"
        r = analyse(body, U)
        r["kernel_total_instr"] = len(body)
        r["utmaldg"] = sum(1 for _, t in body if opcode(t) == "UTMALDG")
        res[f"S{S}{'_lean' if lean else ''}"] = r
    for k in sorted(res):
        r = res[k]
        print(f"{k}: {r['instr_per_column']:.1f} instr / column, fp64 {r['fp64_per_column']:.1f}, local-memory ops per iteration "
              f"{r['local_memory_ops_per_iteration']}, mix {dict(list(r['mix_per_iteration'].items())[:12])}")
    if args.write:
        out = {"kernel_version": args.version, "what": "interior x-loop of k_fused_stage<S, 4, 128>, per thread and column "
               "(cuobjdump -sass of veritas_b200/build/vrt_fused.cu.o, sm_100a); one thread-column = one cell plus the recomputed strip / chunk halo",
               "stages": res}
        json.dump(out, open(os.path.join(ROOT, "profiles", "fused_sass_counts.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
