#!/usr/bin/env python
"""DRAM bytes per launch of each workload's dominant kernel (ncu: dram__bytes_read.sum + dram__bytes_write.sum), written
to profiles/traffic_<workload>.json together with the hash of the kernel's sources -- bench.py reports the figure as
`roofline.traffic` only while that hash still matches (a number measured on an older kernel is reported as stale).
Run on the GPU box:  python tools/measure_traffic.py [workload ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KERNEL = {"fm2b": "fm_", "fm2a": "fm_", "fm1": "fm_", "fm5a": "fm_", "power3": "power_fft", "power4": "power_fft"}
LAUNCHES = {"fm2a": 2}          # kernels per step (the stream path: front kernel + back kernel); their traffic is summed


def main():
    todo = sys.argv[1:] or list(KERNEL)
    for w in todo:
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{KERNEL[w]}",
               "-c", str(2 * LAUNCHES.get(w, 1)), "--csv", sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", "1", "--warmup", "1",
               "--no-e2e", "--no-cpu", "--no-extras"]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout
        lines = [l for l in out.splitlines() if l.startswith('"')]
        rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
        by_id = {}
        for r in rows:
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"].lower()
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            by_id.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[r["Metric Name"]] = v
        if not by_id:
            print(w, "no ncu rows", file=sys.stderr)
            continue
        ids = sorted(by_id, key=int)[-LAUNCHES.get(w, 1):]        # the second step's launches: caches warm like a timed step
        rd = sum(by_id[i].get("dram__bytes_read.sum", 0.0) for i in ids)
        wr = sum(by_id[i].get("dram__bytes_write.sum", 0.0) for i in ids)
        total = rd + wr
        last = {"kernel": " + ".join(by_id[i]["kernel"] for i in ids)}
        rec = {"workload": w, "kernel": last["kernel"], "dram_bytes_per_launch": total,
               "dram_bytes_read": rd, "dram_bytes_write": wr,
               "source_sha16": bench.kernel_source_sha(w),
               "from": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, second launch of bench.py --steps 1 --warmup 1"}
        with open(os.path.join(ROOT, "profiles", f"traffic_{w}.json"), "w") as f:
            json.dump(rec, f, indent=1)
        print(w, rec["kernel"][:60], "%.4f GB" % (total / 1e9))


if __name__ == "__main__":
    main()
