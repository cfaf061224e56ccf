"""Turn gpurun_out/launches.csv + gpurun_out/prof_score.ncu-rep into committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")

rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 20 --warmup 3 --no-cpu`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare shares).\n\n"
            "| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k[:90]}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {100*sum(v)/tot:.1f}% |\n")
import shutil
shutil.copy(os.path.join(go, "launches.csv"), os.path.join(out, f"{tag}_launches.csv"))

raw = subprocess.run(["ncu", "-i", os.path.join(go, "prof_score.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]; units = rr[1]; vals = rr[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
name = vals[h.index("Kernel Name")] if "Kernel Name" in h else "k_score_tc"
got = {}
with open(os.path.join(out, f"{tag}_score_kernel_ncu.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of the score kernel\n\nkernel: `{name}`  (one launch = 1024 graphs, S=50, E=10000)\n\n| metric | value | unit |\n|---|---:|---|\n")
    for w in want:
        if w in h:
            i = h.index(w); got[w] = vals[i]
            f.write(f"| {w} | {vals[i]} | {units[i]} |\n")
try:
    rd = float(got["dram__bytes_read.sum"].replace(",", "")); wr = float(got["dram__bytes_write.sum"].replace(",", ""))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    ur = units[h.index("dram__bytes_read.sum")]; uw = units[h.index("dram__bytes_write.sum")]
    total = rd * scale.get(ur, 1.0) + wr * scale.get(uw, 1.0)
    json.dump({"score_kernel_dram_bytes_per_launch": total, "source": f"profiles/{tag}_score_kernel_ncu.md"}, open(os.path.join(out, "traffic.json"), "w"))
    print("dram bytes per launch", total)
except Exception as ex:
    print("traffic.json not written:", ex)
print(open(os.path.join(out, f"{tag}_launches.md")).read())
print(open(os.path.join(out, f"{tag}_score_kernel_ncu.md")).read())
