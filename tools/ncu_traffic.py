#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an .ncu-rep (ncu --set full capture), merged into
profiles/ncu_traffic.json under a config name: {cfg: {kernel: {"bytes": read + write, "read": .., "write": .., "ms": ..}}}.
bench.py copies the dominant kernel's entry into roofline.traffic and the sum into roofline.traffic_whole_step.
Usage: python tools/ncu_traffic.py <cfg> <rep> [<rep> ...]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    cfg, reps = sys.argv[1], sys.argv[2:]
    try:
        table = json.load(open(OUT))
    except Exception:
        table = {}
    if any(isinstance(v, dict) and "bytes" in v for v in table.values()):
        table = {}                                   # the round-1 flat format
    entry = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw, it = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        for r in rows[2:]:
            name = r[ik].split("(")[0].split("<")[0].replace("void ", "").strip()
            rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
            wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
            ms = float(r[it].replace(",", "")) * TUNIT[units[it]]
            e = entry.setdefault(name, {"bytes": 0.0, "read": 0.0, "write": 0.0, "ms": 0.0, "launches": 0})
            e["bytes"] += rd + wr; e["read"] += rd; e["write"] += wr; e["ms"] += ms; e["launches"] += 1
    for e in entry.values():
        n = e.pop("launches")
        for k in ("bytes", "read", "write", "ms"):
            e[k] = e[k] / n
        e["bytes"], e["read"], e["write"] = int(e["bytes"]), int(e["read"]), int(e["write"])
    table[cfg] = entry
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    print(cfg, {k: round(v["bytes"] / 1e6, 1) for k, v in entry.items()}, "MB; sum", round(sum(v["bytes"] for v in entry.values()) / 1e6, 1))


if __name__ == "__main__":
    main()
