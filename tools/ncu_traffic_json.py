"""Write the `roofline.traffic` record bench.py reads (profiles/rNN_ncu_traffic_<workload>.json) from an
`ncu --set full` report:  ncu -i report.ncu-rep --page raw --csv | python tools/ncu_traffic_json.py <out.json> <workload>
       [kernel-regex] [launch-index]
Picks the launch-index-th (default 0) launch whose name matches the regex (default gemm_sm100_kernel) and records its
dram__bytes_read.sum + dram__bytes_write.sum, duration and grid — measured, not typed in."""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def main():
    out, workload = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else "gemm_sm100_kernel")
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(sys.stdin))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}

    def val(r, key):
        i = idx[key]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    hits = [r for r in rows[hdr + 2:] if len(r) >= len(names) and pat.search(r[idx["Kernel Name"]])]
    r = hits[which]
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    rec = {
        "workload": workload, "kernel": r[idx["Kernel Name"]][:120], "launch_index_among_matches": which,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "time_us": val(r, "gpu__time_duration.sum"), "grid": r[idx["launch__grid_size"]],
        "source": "ncu --set full --clock-control none (tools/profile.sh), one launch",
    }
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
