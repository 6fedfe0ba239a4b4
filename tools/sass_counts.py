"""Static instruction mix of every kernel in libsddc_b200.so from `cuobjdump -sass` (run here, no GPU needed):
    python tools/sass_counts.py > profiles/r02_sass_mix.txt
Counts the opcodes that decide how a kernel uses the machine: fp64 tensor (DMMA) and scalar fp64 (DFMA / DADD / DMUL),
TMA bulk copies (UBLKCP) and their barriers (SYNCS), cp.async (LDGSTS), plain global / shared / local accesses (LDL / STL
= register spills), shuffles and CTA barriers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectraldoublediffusiveconvection_b200", "libsddc_b200.so")
OPS = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "LDL", "STL",
       "SHFL", "BAR", "ATOMG", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+).*LOCAL:(\d+)", ln)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    counts, order, cur = {}, [], None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    arch = re.search(r"arch = (sm_\w+)", sass)
    print("static SASS instruction mix of %s (%s), one line per kernel" % (os.path.basename(LIB), arch.group(1) if arch else "?"))
    print("%-64s %6s %4s %6s %6s | %s" % ("kernel", "instrs", "regs", "smem", "local", " ".join("%6s" % o for o in OPS)))
    tot = collections.Counter()
    for fn in sorted(order, key=lambda f: -counts[f]["_all"]):
        name = demangle(fn)
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("sddc::", "").replace("(int)", "").replace("(bool)", "")
        reg, shm, loc = usage.get(fn, (0, 0, 0))
        print("%-64s %6d %4d %6d %6d | %s" % (name[:64], counts[fn]["_all"], reg, shm, loc,
                                               " ".join("%6d" % counts[fn][o] for o in OPS)))
        tot.update(counts[fn])
    print("%-64s %6d %4s %6s %6s | %s" % ("TOTAL (%d kernels)" % len(order), tot["_all"], "", "", "",
                                           " ".join("%6d" % tot[o] for o in OPS)))
    print("\nNo tcgen05 / UTCMMA instruction appears and none can: the path is fp64 and tcgen05 has no fp64 MMA kind, so the fp64"
          "\ntensor work is DMMA (mma.sync.m8n8k4.f64).  UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier operations.")


if __name__ == "__main__":
    main()
