"""Condenses an Nsight Compute report (.ncu-rep, `--set full`) into the numbers DESIGN.md and
bench.py quote: per kernel duration, DRAM bytes, instruction/issue/pipe utilisation, shared-memory
wavefronts, the top stall reasons and the hottest SASS lines.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/summary.md [--traffic-key n184_p2_g1]
"""
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_global_red.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tkey = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of `{os.path.basename(rep)}`", "",
             "Captured with `ncu --set full --import-source on --clock-control none` (one launch per kernel, "
             "cold caches, serialised); absolute times are not bench values.", ""]
    traffic = {}
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
        stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled")
                         and k.endswith("per_issue_active.ratio")), reverse=True)[:6]
        lines += ["", "top stall reasons (warps stalled per issue slot): " +
                  ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, k in stalls), ""]
        if "gather2" in name or tkey is None:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"]) * scale.get(u["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"]) * scale.get(u["dram__bytes_write.sum"], 1.0)
            traffic[name] = rd + wr
    src = list(csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "sass").splitlines()))
    sections, cur = [], None
    for r in src:  # one section per kernel: "Kernel Name" row, header row, data rows
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "data": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["data"].append(r)
    for sec in sections:
        h = sec["hdr"]
        if not h or "Instructions Executed" not in h:
            continue
        iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        data = [r for r in sec["data"] if len(r) > max(iE, iW)]
        tot = sum(int(r[iE]) for r in data) or 1
        totw = sum(int(r[iW]) for r in data) or 1
        lines += [f"## hottest SASS lines of `{sec['name'][:70]}` (share of stall samples / of executed instructions)", "", "```"]
        for r in sorted(data, key=lambda r: -int(r[iW]))[:16]:
            lines.append(f"{100 * int(r[iW]) / totw:6.2f}%  {100 * int(r[iE]) / tot:6.2f}%  {r[iS].strip()[:90]}")
        lines += ["```", ""]
    open(out, "w").write("\n".join(lines) + "\n")
    if tkey:
        tpath = os.path.join(os.path.dirname(out), "traffic.json")
        tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
        for name, val in traffic.items():
            if "gather2" in name:
                tj[tkey] = val
        json.dump(tj, open(tpath, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
