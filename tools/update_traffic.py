#!/usr/bin/env python
"""Write profiles/fused_traffic.json (the DRAM bytes bench.py quotes as roofline.traffic) from an ncu --set full report of the
fused stage kernel:   python tools/update_traffic.py gpurun_out/prof_fused.ncu-rep v16 profiles/ncu_fused_v16_summary.txt
The second argument is the kernel version the capture was taken on (bench.py only quotes a capture of FUSED_KERNEL_VERSION); the
third names the committed summary of the same report."""
import csv, json, os, re, subprocess, sys
rep, version, source = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, row = r[0], r[1], r[2]          # first profiled launch
get = lambda name: (float(row[hdr.index(name)]), units[hdr.index(name)])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rd, ru = get("dram__bytes_read.sum")
wr, wu = get("dram__bytes_write.sum")
name = row[hdr.index("Kernel Name")]
stage = int(re.search(r"k_fused_stage<\(?(?:int\))?(\d)", name).group(1))
out = {"kernel": re.sub(r"\(int\)", "", name).split("(")[0].replace("void <unnamed>::", ""), "kernel_version": version, "stage": stage,
       "workload": "c3 (65536x4096, one species per launch)", "dram_bytes_read": rd * scale[ru], "dram_bytes_write": wr * scale[wu],
       "source": f"{source} (ncu --set full --clock-control none, one launch)"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fused_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(out)
