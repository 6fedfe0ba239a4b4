"""Instruction mix and stall samples per opcode / per source region of one kernel of an ncu report (run here, no GPU):
    python tools/sass_mix.py gpurun_out/prof.ncu-rep nlin_fft_kernel
"""
import collections
import csv
import io
import subprocess
import sys


def load(rep, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                          "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
    return hdr, data


def main(rep, kern):
    hdr, data = load(rep, kern)
    ix = {h: i for i, h in enumerate(hdr)}
    tot, ninst, st = collections.Counter(), collections.Counter(), collections.Counter()
    samples = inst_total = 0
    stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, c):
        try:
            return int(r[ix[c]])
        except Exception:
            return 0
    for r in data:
        src = r[ix["Source"]].split()
        op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
        op = op.split(".")[0]
        s, n = num(r, "# Samples"), num(r, "Instructions Executed")
        tot[op] += s; ninst[op] += n; samples += s; inst_total += n
        for c in stallcols:
            st[c] += num(r, c)
    print("kernel %s: %d static instructions, %d warp-instructions executed, %d stall samples" % (kern, len(data), inst_total, samples))
    for op, n in ninst.most_common(24):
        print("  %-10s executed %10d (%5.1f%%)   samples %7d (%5.1f%%)" % (op, n, 100.0 * n / max(inst_total, 1), tot[op], 100.0 * tot[op] / max(samples, 1)))
    print("  stall reasons:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / max(samples, 1)) for k, v in st.most_common(10)))
    w = sum(num(r, "L1 Wavefronts Shared") for r in data)
    wi = sum(num(r, "L1 Wavefronts Shared Ideal") for r in data)
    print("  shared-memory wavefronts %d (ideal %d)" % (w, wi))
    static = collections.Counter()
    for r in data:
        src = r[ix["Source"]].split()
        op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
        static[op.split(".")[0]] += 1
    print("  static mix:", ", ".join("%s %d" % kv for kv in static.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])


def phases(rep, kern):
    """Stall samples and executed warp-instructions between consecutive BAR instructions (address order)."""
    hdr, data = load(rep, kern)
    ix = {h: i for i, h in enumerate(hdr)}
    seen, rows = set(), []
    for r in data:
        if r[0] in seen:
            continue
        seen.add(r[0]); rows.append(r)
    rows.sort(key=lambda r: int(r[0], 16))
    stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, c):
        try:
            return int(r[ix[c]])
        except Exception:
            return 0
    seg, segs = dict(s=0, n=0, f=0, st=collections.Counter(), first=None), []
    for r in rows:
        src = r[ix["Source"]]
        seg["s"] += num(r, "# Samples"); seg["n"] += num(r, "Instructions Executed")
        if any(o in src for o in ("DADD", "DMUL", "DFMA")):
            seg["f"] += num(r, "Instructions Executed")
        for c in stallcols:
            seg["st"][c.replace("stall_", "")] += num(r, c)
        if "BAR.SYNC" in src or "BAR.RED" in src:
            segs.append(seg); seg = dict(s=0, n=0, f=0, st=collections.Counter())
    segs.append(seg)
    tot = sum(s["s"] for s in segs)
    for i, s in enumerate(segs):
        print("  segment %2d: samples %6d (%5.1f%%)  warp-insts %9d  fp64 %9d   %s" % (
            i, s["s"], 100.0 * s["s"] / max(tot, 1), s["n"], s["f"],
            ", ".join("%s %d" % kv for kv in s["st"].most_common(5))))


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "phases":
    phases(sys.argv[1], sys.argv[2])
