"""Summarise an ncu report (run here, no GPU needed): per-kernel key metrics -> text, and the dominant kernel's
DRAM traffic per launch of every kernel -> <out>_traffic.json (read by bench.py for roofline.traffic / step_traffic).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_summary.txt
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor (DMMA) pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 (DFMA) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["ncu summary of %s (--set full --clock-control none; per launch)" % os.path.basename(rep), ""]
    traffic = None
    per_kernel = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("== " + name)
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                lines.append("   %-46s %s %s" % (label, r[i], units[i]))
        stalls = {}
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                stalls[h.split("stalled_")[1].split("_per")[0]] = float(r[i])
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        lines.append("   top warp stalls (per issue): " + ", ".join("%s %.2f" % kv for kv in top))
        import re
        if "nlin_fft_kernel" in name or "nlin_fft_staged_kernel" in name:
            traffic = None   # the FFT formulation's kernel is the dominant one when present
        if (re.search(r"synth_kernel<(\(int\))?\d+, *(\(int\))?0>", name) or "synth_ws_kernel" in name or "synth_wsq_kernel" in name
                or "nlin_fft_kernel" in name or "nlin_fft_staged_kernel" in name) and traffic is None:
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            traffic = {"kernel": name, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                       "source": os.path.basename(rep)}
        lines.append("")
        try:
            rd_ = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr_ = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            per_kernel.append({"kernel": name.split("(")[0].replace("void ", ""), "dram_read": rd_, "dram_write": wr_,
                               "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))})
        except Exception:
            pass
    open(out, "w").write("\n".join(lines))
    json.dump({"source": os.path.basename(rep), "kernels": per_kernel,
               "dram_bytes_total": sum(k["dram_read"] + k["dram_write"] for k in per_kernel)},
              open(os.path.splitext(out)[0] + "_traffic.json", "w"), indent=1)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
