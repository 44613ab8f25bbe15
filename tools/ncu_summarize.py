"""Condense `ncu -i report.ncu-rep --page raw --csv` (stdin) into one row per launch:
kernel | time us | dram rd MB | dram wr MB | DRAM GB/s | tensor-pipe active % | SM thr % | DRAM thr % | L2 thr % | regs | grid."""
import csv
import sys

COLS = {
    "time": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor2": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}
UNIT = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rows = list(csv.reader(sys.stdin))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}

    def val(r, key):
        i = idx.get(COLS[key])
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        return v * UNIT.get(units[i], 1.0)

    print("kernel | time us | dram rd MB | dram wr MB | DRAM GB/s | tensor % | SM thr % | DRAM thr % | L2 thr % | regs | grid")
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        t, rd, wr = val(r, "time"), val(r, "rd"), val(r, "wr")
        tens = val(r, "tensor")
        if tens is None:
            tens = val(r, "tensor2")
        gbs = (rd + wr) / t * 1e3 if t and rd is not None and wr is not None else 0.0
        f = lambda v: "n/a" if v is None else f"{v:.2f}"  # noqa: E731
        print(f"{r[idx['Kernel Name']][:60]:60s} | {f(t)} | {f(rd)} | {f(wr)} | {gbs:.0f} | {f(tens)} | {f(val(r, 'sm'))} | "
              f"{f(val(r, 'dram'))} | {f(val(r, 'l2'))} | {f(val(r, 'regs'))} | {f(val(r, 'grid'))}")


if __name__ == "__main__":
    main()
