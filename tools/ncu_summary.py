"""ncu raw-page CSV -> the per-launch tables under profiles/ and profiles/roofline_traffic.json.  Runs on a CPU box.

    # on the GPU box (one step inside cudaProfilerStart/Stop, conv launches only, every metric of the full set):
    ncu --profile-from-start off --set full --clock-control none -k regex:conv_umma --csv --page raw \
        --log-file gpurun_out/conv_family_b1_raw.csv python tools/profile_step.py
    # here:
    python tools/ncu_summary.py --raw gpurun_out/conv_family_b1_raw.csv --out profiles/r2_conv_family_ncu_b1.csv \
        [--traffic-key b1]      # also records the mean DRAM bytes per launch in profiles/roofline_traffic.json

The raw page has one row per launch and one column per metric (second header row = units).  Columns picked, when
present: duration, grid, block, registers, tensor-pipe active %, SM throughput %, L2 throughput %, L2 -> SM bytes, DRAM read / write bytes, DRAM throughput %, warps active %.  roofline_traffic.json carries the hash
of the kernel sources the capture was taken from (tools/sass_summary.py:kernel_sources_sha): bench.py refuses a stale one.
"""
import argparse
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLUMNS = [
    ("time_us", ["gpu__time_duration.sum"], "time"),
    ("grid", ["launch__grid_size"], None),
    ("block", ["launch__block_size"], None),
    ("regs", ["launch__registers_per_thread"], None),
    ("tensor_pipe_pct_active", ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                                "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"], None),
    ("sm_pct", ["sm__throughput.avg.pct_of_peak_sustained_elapsed"], None),
    ("l2_pct", ["lts__throughput.avg.pct_of_peak_sustained_elapsed"], None),
    ("l2_to_sm_MB", ["l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum"], "l2sm"),
    ("dram_read_MB", ["dram__bytes_read.sum"], "bytes"),
    ("dram_write_MB", ["dram__bytes_write.sum"], "bytes"),
    ("dram_pct", ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], None),
    ("warps_active_pct", ["sm__warps_active.avg.pct_of_peak_sustained_active"], None),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3,
              "ms": 1e3, "msecond": 1e3, "second": 1e6}


def read_raw(path):
    """Both CSV shapes of ncu: the raw page (one row per launch, one column per metric, a second header row of units;
    column names may carry a 'UNIT.Section.' prefix) and the --metrics log (one row per launch and metric)."""
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.reader(lines))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[head]
    if "Metric Name" in names:
        iid, ik, im, iu, iv = (names.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
        ib, ig = names.index("Block Size"), names.index("Grid Size")
        launches, metrics, units = {}, [], {}
        for r in rows[head + 1:]:
            if len(r) != len(names):
                continue
            d = launches.setdefault(r[iid], {"Kernel Name": r[ik], "launch__block_size": r[ib].strip("()").split(",")[0],
                                             "launch__grid_size": r[ig].strip("()").split(",")[0]})
            d[r[im]] = r[iv]
            units[r[im]] = r[iu]
            if r[im] not in metrics:
                metrics.append(r[im])
        cols = ["Kernel Name", "launch__block_size", "launch__grid_size"] + [m for m in metrics if m != "launch__grid_size"]
        return cols, [units.get(c, "") for c in cols], [[d.get(c, "nan") for c in cols] for d in launches.values()]
    names = [re.sub(r"^[A-Z_0-9]+\.[A-Za-z]+\.", "", n) for n in names]
    units = rows[head + 1]
    return names, units, [r for r in rows[head + 2:] if len(r) == len(names)]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic-key", default="", help="e.g. b1: record mean DRAM bytes per launch in roofline_traffic.json")
    ap.add_argument("--kernel", default="conv_umma", help="regex: launches that count for the traffic figure")
    a = ap.parse_args()
    names, units, rows = read_raw(a.raw)
    idx = {n: i for i, n in enumerate(names)}
    kcol = idx["Kernel Name"]
    picked = []
    for out_name, cands, kind in COLUMNS:
        col = next((idx[c] for c in cands if c in idx), None)
        picked.append((out_name, col, kind))
    with open(a.out, "w") as f:
        f.write("kernel," + ",".join(n for n, c, _ in picked if c is not None) + "\n")
        total_dram, count = 0.0, 0
        for r in rows:
            kname = re.sub(r"\(.*", "", r[kcol]).replace("gct2::", "").replace("void ", "")
            cells = []
            dram = 0.0
            for out_name, col, kind in picked:
                if col is None:
                    continue
                v = num(r[col])
                u = units[col]
                if kind == "bytes":
                    v = v * UNIT_SCALE.get(u, 1.0)
                    dram += v
                    v /= 1e6
                elif kind == "l2sm":
                    v = v * (32.0 if u == "sector" else UNIT_SCALE.get(u, 1.0)) / 1e6
                elif kind == "time":
                    v = v * UNIT_SCALE.get(u, 1e-3)
                cells.append(f"{v:.3f}")
            f.write(f'"{kname}",' + ",".join(cells) + "\n")
            if re.search(a.kernel, r[kcol]):
                total_dram += dram
                count += 1
    print(f"{len(rows)} launches -> {a.out}")
    if a.traffic_key and count:
        from tools.sass_summary import kernel_sources_sha
        path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        tj = json.load(open(path)) if os.path.exists(path) else {}
        if tj.get("kernel_sha") != kernel_sources_sha():
            tj = {"kernel_sha": kernel_sources_sha()}  # figures of other kernel sources do not mix with these
        tj[f"traffic_bytes_per_launch_{a.traffic_key}"] = total_dram / count
        tj[f"launches_{a.traffic_key}"] = count
        tj.setdefault("sources", {})[a.traffic_key] = os.path.relpath(a.out, ROOT)
        tj["source"] = ", ".join(sorted(tj["sources"].values()))
        tj["how"] = ("sum of dram__bytes_read.sum + dram__bytes_write.sum over the conv_umma launches of one eager training "
                     "step (ncu --profile-from-start off --set full --clock-control none -k regex:conv_umma --csv --page raw "
                     "python tools/profile_step.py [--batch N]), divided by the launch count; tools/ncu_summary.py")
        with open(path, "w") as f:
            json.dump(tj, f, indent=1, sort_keys=True)
        print(f"traffic {a.traffic_key}: {total_dram / count / 1e6:.2f} MB per launch over {count} launches -> {path}")


if __name__ == "__main__":
    main()
