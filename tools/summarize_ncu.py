"""Turn ncu artefacts brought back from the GPU box into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py full  REPORT.ncu-rep OUT.md [--title T]     one `ncu --set full` capture
    python tools/summarize_ncu.py list  LAUNCHES.csv   OUT.md                 the gpu__time_duration launch list
"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "sm__cycles_active.avg",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, out, title):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`, one launch; "
             "cold-cache, serialised — not a bench number).", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        st = [(h, r[i]) for i, h in enumerate(hdr) if "average_warps_issue_stalled" in h and r[i] not in ("", "n/a")]
        st.sort(key=lambda t: -float(t[1]))
        lines += ["", "Warp stall reasons (warps per issue-active cycle): " +
                  ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {float(v):.2f}" for h, v in st[:8]), ""]
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ie, so = h.index("Instructions Executed"), h.index("Source")
        body = []
        for r in src[2:]:
            if r and r[0] == "Kernel Name":
                break
            body.append(r)
        ops = Counter()
        for r in body:
            op = r[so].split()
            if op and op[0].startswith("@"):
                op = op[1:]
            if op:
                ops[op[0].split(".")[0]] += int(r[ie])
        tot = sum(ops.values())
        lines += [f"Executed warp-instructions by opcode (first launch, total {tot}): " +
                  ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(18)), ""]
    open(out, "w").write("\n".join(lines) + "\n")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[kn]].append(float(r[mv]))
    tot = sum(sum(v) for v in agg.values())
    lines = ["# ncu launch list (gpu__time_duration.sum per launch)", "",
             f"Source: `{path.split('/')[-1]}` — `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py`; "
             "per-launch times are cold-cache and serialised, only the SHARES are meaningful.", "",
             "| kernel | launches | total us | share | mean us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k[:110]}` | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v) / 1e3:.2f} |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        t = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else sys.argv[2]
        full(sys.argv[2], sys.argv[3], t)
    else:
        launch_list(sys.argv[2], sys.argv[3])
