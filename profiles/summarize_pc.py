#!/usr/bin/env python
"""Where the warps of the fused rx_fm kernel are: PC-sampling data of one `ncu --set full --import-source on` capture,
grouped by the code regions between the kernel's barriers (ticket fetch | front end | back end | tail).

usage: python profiles/summarize_pc.py gpurun_out/final/prof_fm2b.ncu-rep > profiles/r1_pc_sampling_fm2b.txt
Reads the report through `ncu -i ... --page source --csv --print-source sass` (no GPU needed).
"""
import csv
import io
import subprocess
import sys

STALLS = ["stall_selected", "stall_not_selected", "stall_math", "stall_wait", "stall_barrier", "stall_long_sb",
          "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_membar"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][0] + ":", rows[0][1])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name):
        try:
            return int(r[ix[name]] or 0)
        except (ValueError, IndexError, KeyError):
            return 0

    base = int(data[0][0], 16)
    total = sum(col(r, "# Samples") for r in data)
    inst = sum(col(r, "Instructions Executed") for r in data)
    print(f"samples {total}, warp instructions {inst}")
    regions, cur = [], []
    for r in data:
        cur.append(r)
        if "BAR." in r[1] or "MEMBAR" in r[1]:
            regions.append(cur)
            cur = []
    regions.append(cur)
    print("\nregions between barriers (a sample taken while a warp waits AT a barrier is booked on the instruction after it):")
    for reg in regions:
        n = sum(col(r, "# Samples") for r in reg)
        if n * 200 < total:
            continue
        k = sum(col(r, "Instructions Executed") for r in reg)
        bar = sum(col(r, "stall_barrier") for r in reg)
        print(f"  0x{int(reg[0][0], 16) - base:05x}..0x{int(reg[-1][0], 16) - base:05x}  samples {100 * n / total:5.1f} %  "
              f"(of which waiting at the barrier behind it {100 * bar / total:5.1f} %)  instructions {100 * k / inst:5.1f} %  "
              f"ends with {reg[-1][1].strip()[:40]}")
    print("\nstall reasons, share of all samples:")
    for s in STALLS:
        print(f"  {s[6:]:18s} {100 * sum(col(r, s) for r in data) / total:5.1f} %")
    print("\ntop instructions by samples:")
    for r in sorted(data, key=lambda r: -col(r, "# Samples"))[:12]:
        print(f"  0x{int(r[0], 16) - base:05x}  {100 * col(r, '# Samples') / total:4.1f} %  {r[1].strip()[:60]}")


if __name__ == "__main__":
    main(sys.argv[1])
