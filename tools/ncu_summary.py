"""Turns .ncu-rep captures (ncu --set full) into the small CSV summaries kept under profiles/.

    python tools/ncu_summary.py OUT.csv LABEL=path.ncu-rep [LABEL=path.ncu-rep ...]

Per launch: duration, DRAM bytes, launch geometry, occupancy limits, pipe / issue utilisation and
the warp-stall breakdown; for kernels with block barriers also the PC-sampling totals per barrier
interval (which phase of the pivot the warps spend their time in).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["stall_barrier", "stall_branch_resolving", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math",
          "stall_mio", "stall_lg", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_dispatch",
          "stall_membar"]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    lines = []
    for spec in specs:
        label, rep = spec.split("=", 1)
        raw = page(rep, "raw")
        hdr, units = raw[0], raw[1]
        for li, vals in enumerate(raw[2:]):
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            lines.append([f"## {label} launch {li}"])
            lines.append(["Kernel Name", d.get("Kernel Name", ("", ""))[0]])
            for k in KEYS:
                if k in d:
                    lines.append([k, d[k][0], d[k][1]])
            for h, (v, u) in d.items():
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                    try:
                        if float(v) >= 0.1:
                            lines.append([h, v, u])
                    except ValueError:
                        pass
        src = page(rep, "source")
        # one section per launch: a "Kernel Name" line, the column header, then one row per instruction
        sections, cur = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}
                sections.append(cur)
            elif cur is not None and cur["hdr"] is None and "Source" in r:
                cur["hdr"] = r
            elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
                cur["rows"].append(r)
        for si, sec in enumerate(sections[:1]):
            hdr, data = sec["hdr"], sec["rows"]
            if not hdr or "# Samples" not in hdr:
                continue
            ix = {h: i for i, h in enumerate(hdr)}

            def num(r, k):
                try:
                    return int(r[ix[k]] or 0)
                except ValueError:
                    return 0
            bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[ix["Source"]] or "BARRIER.SYNC" in r[ix["Source"]]]
            if bars:
                lines.append([f"## {label}: PC samples per barrier interval (SASS index range, samples, instructions, top stalls)"])
                prev = 0
                for b in bars + [len(data)]:
                    hi = min(b + 1, len(data))
                    n = sum(num(r, "# Samples") for r in data[prev:hi])
                    ins = sum(num(r, "Instructions Executed") for r in data[prev:hi])
                    st = {k[6:]: sum(num(r, k) for r in data[prev:hi]) for k in STALLS if k in ix}
                    top = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5] if v)
                    lines.append([f"[{prev},{b}]", n, ins, top])
                    prev = b + 1
        lines.append([])
    with open(out_path, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    print("wrote", out_path, len(lines), "lines")


if __name__ == "__main__":
    main()
