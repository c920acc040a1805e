#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` capture: DRAM bytes per launch of the dominant kernel.

    python tools/ncu_traffic.py gpurun_out/r02a_integrate.ncu-rep C3 "C3 at 16 spp" profiles/r02a_ncu_details.txt 16

Reads the report with `ncu -i <rep> --page raw --csv` (runs in the build container: no GPU needed), writes the details page
next to it (`--page details`) and records dram__bytes_read.sum + dram__bytes_write.sum, the duration and the kernel symbol in
profiles/traffic.json (keyed by workload) — bench.py puts exactly these numbers into roofline.traffic, so the line carries a measurement."""
import csv
import io
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep, workload, what, details = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    spp = int(sys.argv[5]) if len(sys.argv) > 5 else None       # samples per pixel of the captured launch
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(header)}
    r = data[-1]                                    # the captured launch
    rec = {"workload": workload, "what": what, "source": details, "spp_captured": spp, "kernel": r[col["Kernel Name"]],
           "dram_bytes_read": to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]),
           "dram_bytes_write": to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]),
           "duration_ms": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[units[col["gpu__time_duration.sum"]]]}
    txt = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True, check=True).stdout
    (ROOT / details).write_text(txt)
    path = ROOT / "profiles" / "traffic.json"
    try:
        allrec = json.loads(path.read_text())
    except Exception:
        allrec = {}
    allrec[workload] = rec                          # one record per workload: the newest capture wins
    path.write_text(json.dumps(allrec, indent=1) + "\n")
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
