#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/ncu_summary.py full  <rep.ncu-rep> <out.md> [workload-key]   # per-kernel table from --set full
  python tools/ncu_summary.py list  <launches.csv> <out.md>                 # launch list, share of the step
With a workload key, `full` also updates profiles/traffic.json (dram bytes per launch, read by bench.py).
"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
           ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers/thread"),
           ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads/instr"),
           ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(rep, out, key=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full --clock-control none: {os.path.basename(rep)}", "",
             "Per-launch metrics (cold-cache, serialised replays: compare shares and bytes, not absolute times).", ""]
    traffic = {}
    for r in data:
        name = short(r[idx["Kernel Name"]])
        lines.append(f"## {name}")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for m, label in METRICS:
            if m in idx:
                lines.append(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        if "dram__bytes_read.sum" in idx:
            t = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            lines.append(f"| **dram traffic per launch** | {t / 1e9:.4f} GB |")
            traffic[name.split("<")[0]] = t
        lines.append("")
    open(out, "w").write("\n".join(lines))
    if key:
        p = os.path.join(ROOT, "profiles", "traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d.setdefault(key, {}).update(traffic)
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print("wrote", out)


def launch_list(csvf, out):
    rows = [r for r in csv.reader(open(csvf)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = {}
    order = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        full_name = r[idx["Kernel Name"]]
        name = short(full_name) if "iso::" in full_name or "signpack" in full_name or "_kernel<" in full_name and "at::" not in full_name \
            else "(torch: synthetic-field setup / tensor fills)"
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)  # -> us
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none: {os.path.basename(csvf)}", "",
             "Every kernel launched by the command, aggregated by name (per-launch times are cold-cache and serialised;",
             "the SHARE column is what is comparable with bench.py's live CUDA-event stage times).", "",
             "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for n in sorted(order, key=lambda k: -agg[k][1]):
        c, t = agg[n]
        lines.append(f"| {n} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        launch_list(sys.argv[2], sys.argv[3])
