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

def raw_page(rep):
    raw = subprocess.run(["ncu", "-i", os.path.join(go, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return rr[0], rr[1], rr[2]
h, units, vals = raw_page("prof_score.ncu-rep")
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
# second capture: the multi-level table kernel at the config-4 shape (tools/kernel_time.py 100 100000 128 lutm)
if os.path.exists(os.path.join(go, "prof_lutm.ncu-rep")):
    h2, u2, v2 = raw_page("prof_lutm.ncu-rep")
    name2 = v2[h2.index("Kernel Name")] if "Kernel Name" in h2 else "k_score_tc"
    with open(os.path.join(out, f"{tag}_lutm_ncu.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of the multi-level table score kernel\n\nkernel: `{name2}`  "
                "(one launch = 128 graphs, S=100, E=100000, E-genes sorted: `python tools/kernel_time.py 100 100000 128 lutm`)\n\n"
                "| metric | value | unit |\n|---|---:|---|\n")
        for w in want + ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
                         "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active"]:
            if w in h2:
                i = h2.index(w)
                f.write(f"| {w} | {v2[i]} | {u2[i]} |\n")
    print(open(os.path.join(out, f"{tag}_lutm_ncu.md")).read())
print(open(os.path.join(out, f"{tag}_launches.md")).read())
print(open(os.path.join(out, f"{tag}_score_kernel_ncu.md")).read())
