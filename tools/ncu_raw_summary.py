"""Tabulate the metrics that matter from an `ncu -i X.ncu-rep --page raw --csv` export, one row
per profiled launch.  Usage: python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel", str),
    ("launch__grid_size", "grid", float),
    ("gpu__time_duration.sum", "time", float),
    ("sm__cycles_elapsed.max", "cycles", float),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%", float),
    ("dram__bytes_read.sum", "dram_rd", float),
    ("dram__bytes_write.sum", "dram_wr", float),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", float),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", float),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc%", float),
    ("lts__t_bytes.sum", "l2_bytes", float),
    ("launch__registers_per_thread", "regs", float),
]


def main():
    rd = list(csv.reader(open(sys.argv[1])))
    hdr, units = rd[0], rd[1]
    pos = {h: i for i, h in enumerate(hdr)}
    names = [c for c in COLS if c[0] in pos]
    print(" | ".join("%s[%s]" % (n[1], units[pos[n[0]]]) for n in names))
    for r in rd[2:]:
        out = []
        for key, label, typ in names:
            v = r[pos[key]]
            if typ is str:
                out.append(v.split("(")[0][-28:])
            else:
                try:
                    out.append("%.4g" % float(v.replace(",", "")))
                except ValueError:
                    out.append(v)
        print(" | ".join(out))


if __name__ == "__main__":
    main()
