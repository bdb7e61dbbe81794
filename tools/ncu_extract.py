"""profiles/ncu_<kind>.json from an ncu raw CSV (`ncu -i X.ncu-rep --page raw --csv > profiles/rNN_*_ncu_raw.csv`).

  python tools/ncu_extract.py profiles/r02_acc_ncu_raw.csv msm_accumulate_kernel msm_accumulate_2_20 [launches_per_unit]

Sums dram__bytes_read + dram__bytes_write over the launches of the named kernel that make up ONE unit of work (one
MSM: its window-group launches; default: all launches in the file) and records per-launch duration and multiply-pipe
utilisation, the git commit and the hash of the kernel sources the profile was taken from.  bench.py prints
roofline.traffic from this file only while that source hash matches the sources it runs.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}


def main():
    path, pattern, kind = sys.argv[1], sys.argv[2], sys.argv[3]
    per_unit = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    launches = [r for r in rows[2:] if len(r) > col["Kernel Name"] and pattern in r[col["Kernel Name"]]]
    if per_unit:
        launches = launches[:per_unit]
    if not launches:
        raise SystemExit(f"no launch of {pattern} in {path}")
    from bench import kernel_source_hash
    try:
        git = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], text=True).strip()
    except Exception:  # noqa: BLE001
        git = None
    rd = [val(r, "dram__bytes_read.sum") for r in launches]
    wr = [val(r, "dram__bytes_write.sum") for r in launches]
    out = {"kernel": pattern, "file": os.path.relpath(path, ROOT), "git": git, "source_hash": kernel_source_hash(kind),
           "launches": len(launches), "dram_bytes": sum(rd) + sum(wr), "dram_read_bytes_per_launch": rd,
           "dram_write_bytes_per_launch": wr,
           "duration_us_per_launch": [val(r, "gpu__time_duration.sum") for r in launches],
           "fmaheavy_pct_per_launch": [val(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed") for r in launches],
           "note": f"dram__bytes_read.sum + dram__bytes_write.sum over the {len(launches)} launch(es) of {pattern} that make up "
                   f"one unit of work, ncu --set full ({os.path.relpath(path, ROOT)})"}
    dst = os.path.join(ROOT, "profiles", f"ncu_{kind}.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(dst, json.dumps({k: out[k] for k in ("launches", "dram_bytes", "duration_us_per_launch", "fmaheavy_pct_per_launch")}))


if __name__ == "__main__":
    main()
