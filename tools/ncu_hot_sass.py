"""Top stall-sampled SASS instructions of each launch in an `ncu --set full --import-source on`
report.  Usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/ncu_hot_sass.py [N]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rows, launches = [], []
for r in csv.reader(sys.stdin):
    if not r:
        continue
    if r[0] == "Kernel Name":
        rows = []
        launches.append((r[1], rows))
    elif r[0] == "Address":
        hdr = r
    elif launches:
        rows.append(r)
for name, rows in launches:
    i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
    tot = sum(int(r[i_s] or 0) for r in rows)
    print("== %s: %d instructions, %d samples" % (name[:60], len(rows), tot))
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][i_s] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        stalls = [(hdr[j], int(r[j])) for j in range(len(hdr)) if hdr[j].startswith("stall_")
                  and "Not Issued" not in hdr[j] and r[j] not in ("", "0")]
        stalls.sort(key=lambda kv: -kv[1])
        print("  #%4d %5.1f%%  %-70s %s" % (i, 100.0 * int(r[i_s]) / max(tot, 1), r[i_src].strip()[:70],
                                        " ".join("%s=%d" % (k[6:], v) for k, v in stalls[:3])))
