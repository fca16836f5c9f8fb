#!/usr/bin/env python
"""Write profiles/fused_traffic.json (the DRAM bytes bench.py quotes as roofline.traffic) from one or more ncu --set full reports
of the fused stage kernel:   python tools/update_traffic.py v17 profiles/ncu_fused_v17_r2c_summary.txt gpurun_out/prof_fused_r2c.ncu-rep [more.ncu-rep]
The first argument is the kernel version the captures were taken on (bench.py only quotes a capture of FUSED_KERNEL_VERSION); the
second names the committed summary of the same reports.  One entry per RK stage found (the first launch of each)."""
import csv, json, os, re, subprocess, sys
version, source, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
stages = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")]
        m = re.search(r"k_fused_stage<\(?(?:int\))?(\d)", name)
        if not m or m.group(1) in stages:
            continue
        get = lambda n: float(row[hdr.index(n)]) * scale[units[hdr.index(n)]]
        stages[m.group(1)] = {"kernel": re.sub(r"\(int\)", "", name).split("(")[0].replace("void <unnamed>::", ""),
                              "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
                              "duration_ms_under_ncu": float(row[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)}
out = {"kernel_version": version, "workload": "c3 (65536x4096, one species per launch)", "stages": stages,
       "source": f"{source} (ncu --set full --clock-control none, one launch per stage)"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fused_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
