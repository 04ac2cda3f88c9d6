#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) and a launch-list CSV into profiles/*.md / traffic.json (run in the build container)."""
import csv, io, json, subprocess, sys, collections

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.sum", "launch__grid_size", "launch__cluster_dim_x",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main(rep, launches_csv, tag, update_traffic=True):
    hdr, units, rows = raw_page(rep)
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu summary {tag}", "", f"source: `{rep}` (ncu --set full --clock-control none, one steady-state launch per kernel)", ""]
    traffic = {}
    for r in rows:
        lines.append(f"## {r[ki][:90]}")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    lines.append(f"- {w} = {r[i]} {units[i]}")
        def val(name):
            for i, h in enumerate(hdr):
                if h == name:
                    v = float(r[i].replace(",", ""))
                    u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return 0.0
        tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        lines.append(f"- dram bytes read+write per launch = {tot / 1e6:.2f} MB")
        key = ("cosched" if "vit_cosched" in r[ki] else "fused" if "vit_fused" in r[ki] else "attn" if "attn_core" in r[ki]
               else "gemm" if "gemm_tn" in r[ki] else "other")
        traffic.setdefault(key, []).append(tot)
        lines.append("")
    # launch list shares
    if launches_csv in ("", "-"):
        open(f"profiles/ncu_summary_{tag}.md", "w").write("\n".join(lines) + "\n")
        if update_traffic:
            pass
        rows2 = None
    else:
        rows2 = [r for r in csv.reader(open(launches_csv)) if len(r) > 10]
    if rows2 is None:
        if not update_traffic:
            return
        rows2 = [["Kernel Name", "Metric Value"]]
    h2 = rows2[0]
    k2, v2 = h2.index("Kernel Name"), h2.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows2[1:]:
        try:
            agg[r[k2][:70]].append(float(r[v2].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    lines += [f"## launch list ({launches_csv}; cold-cache, serialised: compare SHARES)", "", "| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1000:.2f} | {sum(v) / tot * 100:.1f}% |")
    open(f"profiles/ncu_summary_{tag}.md", "w").write("\n".join(lines) + "\n")
    if not update_traffic:
        return
    # merge into traffic.json (bench.py reads the per-launch DRAM bytes of its dominant kernel from here)
    try:
        tj = json.load(open("profiles/traffic.json"))
    except Exception:
        tj = {}
    if traffic.get("gemm"):
        tj["qkv_gemm_dram_bytes_per_launch"] = traffic["gemm"][0]
    if traffic.get("attn"):
        tj["attn_core_dram_bytes_per_launch"] = traffic["attn"][0]
    if traffic.get("cosched"):
        tj["vit_cosched_dram_bytes_per_launch"] = traffic["cosched"][-1]
        tj["vit_cosched_source"] = f"profiles/ncu_summary_{tag}.md"
    if traffic.get("fused"):
        tj["vit_fused_dram_bytes_per_launch"] = traffic["fused"][0]
        tj["vit_fused_source"] = f"profiles/ncu_summary_{tag}.md"
    else:
        tj["source"] = f"profiles/ncu_summary_{tag}.md"
    json.dump(tj, open("profiles/traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], update_traffic=(len(sys.argv) < 5 or sys.argv[4] != "--no-traffic"))
