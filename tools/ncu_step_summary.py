"""Per-kernel table from an `ncu --metrics ... --csv` capture of one step (tools/_run_ncu.sh):
launches, device time, DRAM bytes per launch, DRAM GB/s against the measured HBM peak, tensor-pipe activity.
  python tools/ncu_step_summary.py gpurun_out/ncu_step_r02.csv profiles/ncu_step_r02.md [--gemm-json profiles/ncu_gemm_step_r02.json --hash <csrc_hash> --skip N]
`--skip N` drops the first N launches of every kernel name (warm-up step) before aggregating (N = launches per step)."""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("db200::", "")
    return re.sub(r"\(.*$", "", name)


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for row in r:
        lid = int(row[ix["ID"]])
        rec = launches.setdefault(lid, {"name": short(row[ix["Kernel Name"]]), "grid": row[ix["Grid Size"]],
                                        "block": row[ix["Block Size"]]})
        try:
            rec[row[ix["Metric Name"]]] = float(row[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    return list(launches.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("out")
    ap.add_argument("--title", default="one step")
    ap.add_argument("--gemm-json", default=None)
    ap.add_argument("--hash", default=None)
    ap.add_argument("--workload", default="dalle_example")
    ap.add_argument("--second-half", action="store_true",
                    help="keep only the second half of each kernel's launches (drop the warm-up step)")
    a = ap.parse_args()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm = None
    for k in ("hbm_gbs", "hbm_gbps"):
        if isinstance(peaks.get(k), (int, float)):
            hbm = float(peaks[k])
            break
    if hbm is None:
        for k, v in peaks.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                hbm = float(v)
                break
    L = load(a.csv)
    by = collections.OrderedDict()
    for rec in L:
        by.setdefault(rec["name"], []).append(rec)
    if a.second_half:
        by = collections.OrderedDict((k, v[len(v) // 2:]) for k, v in by.items())
    T = "gpu__time_duration.sum"
    total = sum(r.get(T, 0.0) for v in by.values() for r in v)
    rows = []
    for name, v in by.items():
        t = sum(r.get(T, 0.0) for r in v)
        rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in v)
        wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in v)
        tp = [r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") for r in v]
        tp = [x for x in tp if x is not None]
        # time-weighted tensor-pipe activity
        tpw = (sum(r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * r.get(T, 0.0)
                   for r in v) / t) if t > 0 and tp else None
        rows.append((t, name, len(v), rd, wr, tpw))
    rows.sort(reverse=True)
    with open(a.out, "w") as f:
        f.write(f"# ncu per-kernel metrics — {a.title}\n\n")
        f.write(f"Source: `{os.path.basename(a.csv)}` (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                "dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed "
                "--clock-control none`; every launch replayed serialised with cold caches: compare shares and bytes, "
                "not absolute times).\n")
        if hbm:
            f.write(f"HBM % = (read+write bytes / time) / {hbm:.0f} GB/s (MEASURED_PEAKS.json).\n")
        f.write("\n| kernel | launches | time us | share | us/launch | DRAM MB/launch (rd+wr) | GB/s | HBM % | tensor-pipe % |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for t, name, n, rd, wr, tpw in rows:
            gbps = (rd + wr) / t if t > 0 else 0.0        # bytes/ns = GB/s
            f.write(f"| `{name}` | {n} | {t / 1e3:.1f} | {100 * t / total:.1f}% | {t / n / 1e3:.1f} | "
                    f"{rd / n / 1e6:.2f}+{wr / n / 1e6:.2f} | {gbps:.0f} | "
                    f"{(100 * gbps / hbm) if hbm else float('nan'):.1f} | "
                    f"{'' if tpw is None else f'{tpw:.1f}'} |\n")
        f.write(f"\nTotal device time of the listed launches: {total / 1e6:.3f} ms.\n")
    if a.gemm_json:
        g = [r for k, v in by.items() if k.startswith("gemm_tc") for r in v]
        rec = {"source": f"tools/_run_ncu.sh -> {os.path.basename(a.csv)} (all gemm_tc launches of one step)",
               "workload": a.workload, "csrc_hash": a.hash, "launches": len(g),
               "dram_read_bytes": sum(r.get("dram__bytes_read.sum", 0.0) for r in g),
               "dram_write_bytes": sum(r.get("dram__bytes_write.sum", 0.0) for r in g),
               "ncu_time_ms": sum(r.get(T, 0.0) for r in g) / 1e6}
        rec["dram_bytes_per_launch"] = (rec["dram_read_bytes"] + rec["dram_write_bytes"]) / max(len(g), 1)
        with open(a.gemm_json, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
