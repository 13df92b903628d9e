#!/usr/bin/env python3
"""Static instruction mix of the front-end STORE loop of fm_fused_kernel<P,SPEC> (the loop that holds
the PCM STS.U16 and the 256-bit stream load).  The kernel is integer-issue bound, so this count is the
first thing to look at before spending GPU time:  tools/sass_loop.py fm_kernels.o [P SPEC [WIDTH]]"""
import collections
import re
import subprocess
import sys

FMA = ("IMAD", "FMUL", "FFMA", "FADD", "HFMA2", "IMUL")
ALU = ("SHF", "LEA", "LOP3", "VIADD", "ISETP", "IADD3", "PRMT", "MOV", "SEL", "IABS", "VIMNMX", "VIADDMNMX",
       "SGXT", "IADD", "BMSK", "FLO", "POPC", "PLOP3", "VABSDIFF", "I2FP", "FMNMX", "FSEL", "FSETP")


def main():
    obj = sys.argv[1]
    P, spec, width = (sys.argv[2:5] + ["128"])[:3] if len(sys.argv) > 3 else ("3", "1", "128")
    fun = "_ZN3rxb15fm_fused_kernelILi%sELi%sELi%sEEEvNS_5FmDevENS_6FmCallE" % (P, spec, width)
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
    ins = []
    for line in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    addr = {a: i for i, (a, _, _) in enumerate(ins)}
    loops = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) in addr and int(m.group(1), 16) < a:
                j = addr[int(m.group(1), 16)]
                body = ins[j:i + 1]
                ops = [o for _, o, _ in body]
                if 1 <= sum(1 for o in ops if o.startswith("LDG") and "256" in o) <= 3 and not any(o.startswith("BAR") for o in ops):
                    loops.append(body)
    if not loops:
        print("no stream loop found")
        return
    for body in loops:
        report(body)


def report(best):
    kind = "STORE (owned segment)" if any(o.startswith("STS.U16") for _, o, _ in best) else "replay (halo)"
    # rare blocks inside the loop: the chunk-start bookkeeping (an inner backward loop), the out-of-line atan2
    # of a chunk's first sample (CALL) and the generic division behind a non-positive divisor (I2F.RP .. BSYNC)
    rare = set()
    baddr = {a: i for i, (a, _, _) in enumerate(best)}
    for i, (a, op, rest) in enumerate(best[:-1]):
        m = re.search(r"0x([0-9a-f]+)", rest) if op.startswith("BRA") else None
        if m and int(m.group(1), 16) in baddr and int(m.group(1), 16) < a:
            rare.update(range(baddr[int(m.group(1), 16)] - 1, i + 1))
        if op.startswith("CALL"):
            rare.update(range(i - 1, i + 4))
        if op.startswith("I2F.RP"):
            j = i
            while j > 0 and not best[j][1].startswith("LOP3.LUT"):
                j -= 1
            k2 = i
            while k2 < len(best) - 1 and not best[k2][1].startswith("BSYNC"):
                k2 += 1
            rare.update(range(j, k2 + 1))
    total = len(best)
    spills = sum(1 for _, o, _ in best if o.startswith(("LDL", "STL")))
    best = [x for i, x in enumerate(best) if i not in rare]
    c = collections.Counter()
    for _, op, _ in best:
        base = op.split(".")[0]
        c["fma" if base in FMA else "alu" if base in ALU else "other"] += 1
    print("%s loop 0x%x..0x%x: common path %d of %d instructions  fma %d  alu %d  other %d  local-memory ops %d"
          % (kind, best[0][0], best[-1][0], len(best), total, c["fma"], c["alu"], c["other"], spills))
    d = collections.Counter(op.split(".")[0] + (".MOV" if ".MOV" in op else "") for _, op, _ in best)
    print("  " + "  ".join("%s %d" % kv for kv in d.most_common(14)))


if __name__ == "__main__":
    main()
