"""Summarise an `ncu --set full` report: one line per kernel launch (duration, DRAM bytes, tensor-pipe and L2
activity, registers), and optionally the per-call-site DRAM traffic JSON that bench.py's `roofline.traffic` reads.

    python tools/ncu_summary.py profiles/r1_prof_tc_gemm.ncu-rep [--match seg_gemm_tc] [--sites a,b,c --json out.json]

Runs `ncu -i <rep> --page raw --csv` (works in the build container: reading a report needs no GPU).
`--sites` names the launches in order (one training step captured in launch order); launches of the same site add up."""
import argparse
import csv
import io
import json
import subprocess
import sys

WANT = {                                   # exact column names of `--page raw`
    "l2_rd": "lts__t_bytes_op_read.sum" ,
    "dur": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,       # -> bytes
        "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}  # -> us


def find(header, key):
    return header.index(key) if key in header else None


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--match", default="", help="substring of the kernel name to keep")
    ap.add_argument("--sites", default="", help="comma-separated call-site names of the kept launches, in order")
    ap.add_argument("--json", default="", help="write {dram_bytes_per_step: {site: bytes}} here")
    a = ap.parse_args()
    if a.report.endswith(".csv"):          # an exported `ncu -i rep --page raw --csv` (reports above 64 MiB do not travel)
        rows = list(csv.reader(open(a.report)))
    else:
        raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True)
        if raw.returncode != 0:
            sys.exit(raw.stderr[-2000:])
        rows = list(csv.reader(io.StringIO(raw.stdout)))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {k: find(header, v) for k, v in WANT.items()}
    name_i, grid_i, block_i = header.index("Kernel Name"), header.index("Grid Size"), header.index("Block Size")
    sites = [s for s in a.sites.split(",") if s]
    traffic, k = {}, 0
    for r in body:
        if a.match and a.match not in r[name_i]:
            continue

        def val(key):
            i = col[key]
            if i is None:
                return float("nan")
            return fnum(r[i]) * UNIT.get(units[i], 1.0)

        site = sites[k] if k < len(sites) else r[name_i].split("(")[0][-28:]
        rd, wr = val("dram_rd"), val("dram_wr")
        kname = r[name_i].split("(")[0].split("::")[-1][:26]
        gbs = (rd + wr) / max(val('dur'), 1e-9) / 1e3          # bytes / us = MB/s -> GB/s
        print(f"{site:18s} {kname:26s} grid={r[grid_i]:>12s} block={r[block_i]:>11s} dur={val('dur'):7.2f} us  "
              f"dram_rd={rd / 1e6:7.2f} MB dram_wr={wr / 1e6:6.2f} MB ({gbs:6.0f} GB/s)  tensor_pipe_active={val('tensor_pct'):5.1f}%  "
              f"lts_throughput={val('lts_pct'):5.2f}%  regs={int(val('regs')) if val('regs') == val('regs') else -1}")
        traffic[site] = traffic.get(site, 0.0) + rd + wr
        k += 1
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"source": f"{a.report} (ncu --set full)", "dram_bytes_per_step": traffic}, f, indent=1)


if __name__ == "__main__":
    main()
