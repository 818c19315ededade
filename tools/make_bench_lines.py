"""profiles/<tag>_bench_lines.md from the bench JSON lines a GPU visit left under gpurun_out/.
    python tools/make_bench_lines.py TAG OUT.md  [title]      (reads gpurun_out/TAG_bench_*.json, TAG_ref_*.json)"""
import glob, json, os, sys

tag, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else f"Bench lines of GPU visit {tag}"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows, extra = [], []
for path in sorted(glob.glob(os.path.join(root, "gpurun_out", f"{tag}_bench_*.json"))):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        continue
    name = os.path.basename(path)[len(tag) + 7:-5]
    ps, ro, r, e = d.get("per_step_launch"), d.get("rollout"), d.get("roofline"), d.get("e2e")
    if ps is None:  # c1 / c5r lines
        extra.append((name, d))
        continue
    f = lambda v, n=3: "-" if v is None else f"{v:.{n}f}"  # noqa: E731
    rows.append("| {} ({} GPU) | {:.3e} | {} | {} | {:.3e} | {} ({} streams) | {} | {} | {} | {:.3e} | {} | {} / {} |".format(
        d["config"]["workload"].split(" (")[0] + f" [{name}]", d["n_gpus"], d["value"], f(d["ms_per_step"] * 1e3, 2), f(r["frac"]),
        ps["value"], f(ps["frac"]), ps["streams"], f(ps["frac_single_stream"]),
        f(ro["block"]["frac"]) + f" (K={ro['steps_per_launch']})" if ro else "-", f(ro["philox"]["frac"]) if ro else "-",
        e["value"], f(e.get("frac_of_pcie_ceiling")), d["clocks"]["sm_mhz"], ",".join(d["clocks"]["reasons"]) or "none"))
    if d.get("acting_c5r"):
        extra.append((name + ".acting_c5r", d["acting_c5r"]))
    if d["config"].get("one_gpu_same_workload"):
        extra.append((name + ".one_gpu_same_workload", d["config"]["one_gpu_same_workload"]))
    for k in ("cpu_baseline", "cpu_baseline_literal"):
        if d.get(k):
            extra.append((name + "." + k, d[k]))
lines = [f"# {title}", "",
         "`value` = headline path (`uavca_rollout`, action block, one stream unless the line says otherwise); `per-step` = one launch per "
         "step; `frac` = SURVEY 8d algorithmic bytes / time / 6,515.7 GB/s.", "",
         "| workload | value (UAV env-steps/s) | us/step | frac | per-step value | per-step frac | per-step frac, 1 stream | rollout block frac | rollout Philox frac | e2e | e2e / PCIe ceiling | SM MHz / reasons |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"] + rows + [""]
for path in sorted(glob.glob(os.path.join(root, "gpurun_out", f"{tag}_ref_*.json"))):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        extra.append((os.path.basename(path)[len(tag) + 1:-5], {k: d.get(k) for k in ("value", "cpu_baseline", "cpu_baseline_literal", "config")}))
    except Exception:
        pass
if extra:
    lines += ["Other objects of the same visit:", "", "```"]
    lines += [f"{n}: {json.dumps(v)}" for n, v in extra]
    lines += ["```", ""]
open(out, "w").write("\n".join(lines))
print("wrote", out, len(rows), "rows")
