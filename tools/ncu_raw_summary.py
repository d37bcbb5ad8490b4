"""Per-launch table from an `ncu --set full` report (read here, no GPU needed):
  python tools/ncu_raw_summary.py gpurun_out/prof_x.ncu-rep profiles/ncu_x_r02.md "title" """
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0), ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]


def short(name):
    name = re.sub(r"^void\s+", "", name).replace("db200::", "")
    return re.sub(r"\(.*$", "", name)


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`, every launch replayed "
                "~40 times with flushed caches: durations are cold-cache, compare ratios).\n\n")
        f.write("| # | kernel | grid x block | " + " | ".join(f"{lab} ({units[ix[m]]})" if m in ix and lab in ("us",) else lab
                                                              for m, lab, _ in COLS if m in ix) + " |\n")
        f.write("|---|---|---|" + "---|" * sum(1 for m, _, _ in COLS if m in ix) + "\n")
        for n, r in enumerate(rows[2:]):
            vals = []
            for m, lab, _ in COLS:
                if m not in ix:
                    continue
                v = r[ix[m]].replace(",", "")
                try:
                    x = float(v)
                    if units[ix[m]] == "byte":
                        x /= 1e6
                    elif units[ix[m]] == "Kbyte":
                        x /= 1e3
                    elif units[ix[m]] == "Gbyte":
                        x *= 1e3
                    elif units[ix[m]] in ("ns", "nsecond"):
                        x /= 1e3
                    elif units[ix[m]] in ("ms", "msecond"):
                        x *= 1e3
                    vals.append(f"{x:.1f}" if x < 1000 else f"{x:.0f}")
                except ValueError:
                    vals.append(v)
            f.write(f"| {n} | `{short(r[ix['Kernel Name']])}` | {r[ix['Grid Size']]} x {r[ix['Block Size']]} | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    main()
