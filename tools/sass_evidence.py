#!/usr/bin/env python
"""Per-kernel counts of the instructions that prove the data paths (bulk TMA, mbarrier, cp.async, 256-bit loads, FP32 forms,
no tensor-core instruction) from the shipped library:  cuobjdump -sass rx_tools_b200/librxb200.so | python tools/sass_evidence.py"""
import collections
import re
import sys

PATS = {'UBLKCP': r'\bUBLKCP', 'SYNCS': r'\bSYNCS', 'LDG.256': r'LDG\.E\.[A-Z0-9.]*256', 'LDGSTS': r'\bLDGSTS', 'LDS': r'\bLDS', 'STS': r'\bSTS',
        'IMAD.HI': r'IMAD\.HI', 'FFMA': r'\bFFMA', 'FADD': r'\bFADD', 'BAR': r'\bBAR\.', 'HMMA/UTCMMA': r'HMMA|UTCMMA|UTCHMMA|TCGEN'}
print("SASS evidence from `cuobjdump -sass rx_tools_b200/librxb200.so` (sm_100a cubins of the shipped library, final round-2 build:")
print("the one the bench lines and ncu summaries of profiles/r2_final_session.log were taken on).  LDGSTS = cp.async (the back")
print("kernel's window fills), UBLKCP = bulk TMA copy (rx_power hop buffers), SYNCS = mbarrier operations.\n")
cur, cnt, samples = None, collections.OrderedDict(), {}
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None or re.match(r'\s*/\* 0x', line):
        continue
    if re.search(r'/\*[0-9a-f]{4,5}\*/', line):
        cnt[cur]['instructions'] += 1
        for k, p in PATS.items():
            if re.search(p, line):
                cnt[cur][k] += 1
                if k in ('UBLKCP', 'LDGSTS', 'SYNCS') and (cur, k) not in samples:
                    samples[(cur, k)] = re.sub(r'\s*/\* 0x.*', '', line).strip()
for f, c in cnt.items():
    if not re.search(r'fm_|power_|sdr_', f):
        continue
    print(f)
    print("    instructions %d  " % c['instructions'] + "  ".join("%s=%d" % (k, c[k]) for k in PATS))
print("\nsample lines:")
for (f, k), l in list(samples.items())[:12]:
    print("  %s  [%s]\n      %s" % (k, f[:70], l))
