"""DRAM traffic per training step by kernel class from the per-launch ncu pass
(`tools/profile_round.sh` -> gpurun_out/step_metrics_raw.csv).  Writes profiles/r2_step_traffic.json,
which bench.py reports as roofline.traffic.  Usage: python tools/step_traffic.py raw.csv out.json"""
import csv
import json
import sys

rd = list(csv.reader(open(sys.argv[1])))
hdr, units = rd[0], rd[1]
pos = {h: i for i, h in enumerate(hdr)}


def val(r, key):
    v = float(r[pos[key]].replace(",", ""))
    u = units[pos[key]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


cls = {}
for r in rd[2:]:
    name = r[pos["Kernel Name"]].split("(")[0].replace("rsu::", "").replace("void ", "")
    name = name.split("<")[0]
    c = cls.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_ns": 0.0})
    c["launches"] += 1
    c["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    t = float(r[pos["gpu__time_duration.sum"]].replace(",", ""))
    c["time_ns"] += t * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(
        units[pos["gpu__time_duration.sum"]], 1)
conv = sum(v["dram_bytes"] for k, v in cls.items() if k in ("conv_gemm_kernel", "conv_gemm2_kernel", "conv_halo_kernel", "conv_halo2_kernel", "first_conv_kernel"))
out = {"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every launch of one training step "
                 "(tools/profile_round.sh, flagship config, batch 32)",
       "conv_class_dram_bytes_per_step": conv, "by_kernel": cls}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: (v["launches"], round(v["dram_bytes"] / 1e9, 2), round(v["time_ns"] / 1e6, 3)) for k, v in cls.items()}, indent=0))
