"""Turns ncu output brought back from the GPU box into the tracked evidence under profiles/ (no GPU needed):
  python tools/ncu_summary.py launches <launches.csv> <out.txt>          per-kernel launch counts, device time and shares
  python tools/ncu_summary.py traffic <raw.csv> <key> <kernel regex>    dram bytes / duration / pipes of the matching launches ->
                                                                         profiles/roofline_traffic.json[key] (bench.py reads it)
<raw.csv> is `ncu -i prof.ncu-rep --page raw --csv`; <launches.csv> the `--metrics gpu__time_duration.sum --csv` log."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def launches(path, out):
    agg = {}
    for r in rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"{sum(a[0] for a in agg.values())} launches, {tot:.3f} ms under ncu (cold caches, serialised: compare shares)\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"  {name[:70]:70s} {n:6d} {ms:10.3f} ms {100 * ms / tot:6.1f} %\n")
    print(open(out).read())


def traffic(path, key, pattern, write=True):
    sel = {}
    table = rows(path)
    if table and "Metric Name" not in table[0]:
        # `--page raw` is wide: one row per launch, one column per metric, and the first data row holds the units
        units = table[0]
        for r in table[1:]:
            if re.search(pattern, r["Kernel Name"]):
                sel[r["ID"]] = {k: (v.replace(",", ""), units.get(k, "")) for k, v in r.items() if v not in (None, "")}
    else:
        for r in table:
            if re.search(pattern, r["Kernel Name"]):
                sel.setdefault(r["ID"], {})[r["Metric Name"]] = (r["Metric Value"].replace(",", ""), r.get("Metric Unit", ""))
    if not sel:
        raise SystemExit(f"no launch matches {pattern}")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def get(d, name, units=None):
        if name not in d:
            return None
        v, u = d[name]
        try:
            return float(v) * (units or {}).get(u, 1.0)
        except ValueError:
            return None
    out = []
    for lid, d in sel.items():
        rd, wr = get(d, "dram__bytes_read.sum", scale), get(d, "dram__bytes_write.sum", scale)
        out.append({"launch": lid, "dram_bytes": (rd or 0) + (wr or 0), "dram_read": rd, "dram_write": wr,
                    "duration_ms": get(d, "gpu__time_duration.sum", {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}),
                    "registers": get(d, "launch__registers_per_thread"),
                    "pipe_fma_cycles_active_pct": get(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                    "pipe_xu_pct": get(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                    "pipe_alu_pct": get(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": get(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "inst_executed": get(d, "smsp__inst_executed.sum"),
                    "thread_inst_per_inst": get(d, "smsp__thread_inst_executed_per_inst_executed.ratio")})
    best = max(out, key=lambda o: o["duration_ms"] or 0)
    if not write:
        return best
    tab_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    tab = json.load(open(tab_path)) if os.path.isfile(tab_path) else {}
    tab[key] = dict(best, source=f"ncu --set full, {os.path.basename(path)}, launch {best['launch']} of {len(out)} matching {pattern}")
    with open(tab_path, "w") as f:
        json.dump(tab, f, indent=1)
    print(json.dumps(tab[key], indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
