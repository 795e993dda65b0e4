"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r1
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launch_list(name=None, what=None):
    name = name or tag
    src = OUT / f"launches_{name}.csv"
    if not src.exists():
        return
    lines = [l for l in src.read_text().splitlines(True) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        kname = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        a = agg.setdefault(kname, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    tot = sum(a[1] for a in agg.values())
    with open(ROOT / "profiles" / f"launches_{name}.txt", "w") as f:
        f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none\n"
                f"# {what or 'python profiles/run_profile.py --steps 4   (window of 4 steady-state steps, 4096 members, 500 basins, one member group)'}\n"
                "# per-launch times are cold-cache and serialised: compare the SHARES with roofline.kernels_ms\n")
        f.write(f"{'kernel':34s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'max_us':>9s} {'share':>6s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:34s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {a[2]:9.1f} {a[1] / tot:6.3f}\n")
        f.write(f"{'TOTAL':34s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}\n")


def full_reports():
    reps = sorted(OUT.glob(f"{tag}_k_*.ncu-rep")) + sorted(OUT.glob(f"{tag}_k_*.raw.csv"))
    for rep in reps:
        if rep.suffix == ".csv":  # converted on the GPU box (`ncu -i REP --page raw --csv`)
            raw = rep.read_text()
            rep = rep.with_name(rep.name.replace(".raw.csv", ""))
        else:
            raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True,
                                 text=True).stdout
        rows = list(csv.reader([ln for ln in raw.splitlines() if not ln.startswith("==")]))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, zip(units, vals)))
        with open(ROOT / "profiles" / f"{rep.stem if rep.suffix == '.ncu-rep' else rep.name}.txt", "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {d.get('Kernel Name', ('', '?'))[1]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:90s} {d[k][1]:>16s} {d[k][0]}\n")
            rd = float(d["dram__bytes_read.sum"][1].replace(",", "")) if "dram__bytes_read.sum" in d else 0
            wr = float(d["dram__bytes_write.sum"][1].replace(",", "")) if "dram__bytes_write.sum" in d else 0
            f.write(f"# dram traffic (read+write): {rd + wr:.3f} {d['dram__bytes_read.sum'][0]}\n")


def summary_table():
    """Achieved DRAM bandwidth per kernel (ncu dram bytes / duration of one full launch) against
    the box-measured copy peak and the nominal 8 TB/s (BASELINE.json: 'RHS HBM GB/s')."""
    import json
    import re
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6542.1
    rows = []
    for f in sorted((ROOT / "profiles").glob(f"{tag}_k_*.txt")):
        txt = f.read_text()
        name = re.search(r"one launch of (?:void )?([^(]+)", txt).group(1).strip()
        us = float(re.search(r"gpu__time_duration.sum\s+([\d.,]+)\s+us", txt).group(1).replace(",", ""))
        m = re.search(r"# dram traffic \(read\+write\): ([\d.]+) (\w+)", txt)
        mb = float(m.group(1)) * {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3}[m.group(2)]
        gbs = mb * 1e6 / (us * 1e-6) / 1e9
        rows.append((name, us, mb, gbs))
    with open(ROOT / "profiles" / f"summary_{tag}.md", "w") as out:
        out.write(f"# ncu --set full, one full launch each (4096 members, 500 basins), round {tag}\n\n"
                  f"Peak: {peak:.0f} GB/s measured copy bandwidth (MEASURED_PEAKS.json); 8000 GB/s nominal.\n\n"
                  "| kernel | duration (us) | DRAM read+write (MB) | GB/s | % of measured peak | % of 8 TB/s |\n"
                  "|---|---|---|---|---|---|\n")
        for name, us, mb, gbs in rows:
            out.write(f"| `{name}` | {us:.1f} | {mb:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | {100 * gbs / 8000:.0f} |\n")
        for solve in ("k_newton_linear_ent", "k_newton_linear_lane"):
          newton = [r for r in rows if any(k in r[0] for k in ("k_basin<1", "k_links<1", "k_assemble_linear",
                                                                 solve, "k_newton_update"))]
          if len(newton) == 5:
              mb = sum(r[2] for r in newton)
              us = sum(r[1] for r in newton)
              # SURVEY 8(d) for the 500-basin network: B_rhs + B_jac + B_solve per member and iteration
              alg_kb = (48 * 500 + 8 * 1086 + 16 * 10 + 32 * 500 + 8 * 10 + 8 * 2138 + 8 * 1621 + 16 * 510) / 1e3
              out.write(f"\nOne Newton iteration of 4096 members (basin, links, assembly, `{solve}`, update): {mb:.0f} MB of DRAM "
                        f"traffic in {us:.0f} us = {mb * 1e3 / 4096:.0f} KB per member against {alg_kb:.0f} KB of algorithmic "
                        f"bytes (SURVEY 8(d): B_rhs + B_jac + B_solve): {mb * 1e3 / 4096 / alg_kb:.1f}x.  Every phase boundary is "
                        "a kernel boundary in the member-fastest layout: level / factor and their partials, the J values "
                        "(written by k_links, read by k_assemble_linear and by k_newton_update), the factor slots and the "
                        "reduced right-hand side round-trip through HBM, and the full-state vectors z and du are read by "
                        "three kernels.\n")


if tag == "r1":
    launch_list()
else:
    launch_list(tag, "python bench.py --ncu-range --steps 4 --spinup 48 --no-extra --no-cpu-baseline --no-profile --groups 1"
                     "   (the device-resident timed leg: one window of 4 steps, 4096 members, 500 basins)")
    launch_list(tag + "tree", "python bench.py --ncu-range --workload tree --steps 1 --spinup 3 --warmup 1 --no-extra "
                              "--no-cpu-baseline --no-profile --groups 1   (one step of the 100k-basin tree, 256 members, "
                              "still in its start-up transient)")
full_reports()
summary_table()
print("written:", sorted(p.name for p in (ROOT / "profiles").glob(f"*{tag}*")))
