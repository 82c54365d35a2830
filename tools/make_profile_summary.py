"""Writes the judged profile summaries under profiles/ from the scratch ncu outputs in gpurun_out/.

    python tools/make_profile_summary.py <tag> <launches.csv> <eval.ncu-rep> [<pp_scan.ncu-rep>
                                         [<eval_prune.ncu-rep>]]
"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, eval_rep = sys.argv[1], sys.argv[2], sys.argv[3]
pp_rep = sys.argv[4] if len(sys.argv) > 4 else None
prune_rep = sys.argv[5] if len(sys.argv) > 5 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def launch_table(path):
    rows = list(csv.reader(open(path)))
    for k, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, k + 1
            break
    ix = {h: i for i, h in enumerate(hdr)}
    seq = []
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1.0)
        seq.append((r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]], v))
    return seq


seq = launch_table(launches)
# the device-resident steps are the launches whose eval grid is the full 100000-scenario batch
big = [i for i, (n, g, v) in enumerate(seq) if n.startswith("eval_kernel") and g.startswith("(100000")]
lines = ["# ncu launch list, %s" % tag, "",
         "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
         "--steps 2 --warmup 1 --quick` (cold-cache, serialised "
         "launches: compare shares, not absolutes).", "",
         "%d launches captured; the device-resident steps (eval grid = 100000 CTAs):" % len(seq), "",
         "| step | kernel | grid | ms | share of step |", "|---|---|---|---|---|"]
for si, i in enumerate(big):
    grp = seq[i - 3:i + 2]   # pp_scan, pp_finish, sample_warp, eval, select
    tot = sum(v for _, _, v in grp)
    for n, g, v in grp:
        lines.append("| %d | %s | %s | %.4f | %.1f %% |" % (si, n, g, v, 100 * v / tot))
lines += ["", "Other launches: LUT build, clearance map + distance transform (3), peak microbenchmarks."]
open(os.path.join(out_dir, "%s_launches.md" % tag), "w").write("\n".join(lines) + "\n")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "idc__request_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def kernel_summary(rep, title, kern_sub, name, stages=False):
    m = raw_metrics(rep)
    src_csv = os.path.join(ROOT, "gpurun_out", name + "_src.csv")
    with open(src_csv, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                               text=True).stdout)
    by_line = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), src_csv, kern_sub,
                              os.environ.get("F1L_PROFILE_SO", os.path.join(ROOT, "f1tenth_planning_b200", "lib", "libf1l.so")), "30"],
                             capture_output=True, text=True).stdout
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), src_csv],
                         capture_output=True, text=True).stdout
    lines = ["# %s" % title, "", "`ncu --set full --clock-control none --import-source on`, one launch; "
             "read with `ncu -i ... --page raw|source --csv` and tools/ncu_by_line.py.", "", "## metrics", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in WANT:
        if k in m:
            lines.append("| %s | %s | %s |" % (k, m[k][1], m[k][0]))
    if stages:
        reg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), src_csv, kern_sub],
                             capture_output=True, text=True).stdout
        lines += ["", "## stages (tools/ncu_regions.py: outermost call site of every SASS instruction)", "",
                  reg.strip()]
    lines += ["", "## instruction mix (warp-instructions executed)", "", "```", mix.strip(), "```", "",
              "## stall samples by CUDA source line", "", "```", by_line.strip(), "```"]
    open(os.path.join(out_dir, "%s.md" % name), "w").write("\n".join(lines) + "\n")


def kernel_of(rep):
    """mangled-name substring of the (single) kernel in a report"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[2][rows[0].index("Kernel Name")]
    return name


ek = kernel_of(eval_rep)
targs = ek[ek.index("<") + 1:ek.index(">")].replace(" ", "").split(",")
sub = "eval_kernelILi%sELi%sELi%sELi%sELi%s" % tuple(targs)
kernel_summary(eval_rep, "%s (K3+K4), bench workload (every window segment tested), %s" % (ek.split("(")[0].replace("void ", ""), tag),
               sub, "%s_eval_kernel" % tag, stages=True)
if pp_rep:
    kernel_summary(pp_rep, "pp_scan_kernel (K1 scan), 10^5 poses x 1999 segments, %s" % tag, "pp_scan_kernel",
                   "%s_pp_scan_kernel" % tag)
if prune_rep:
    kernel_summary(prune_rep, "%s (K3+K4), bench workload with prune_window = 1, %s" % (ek.split("(")[0].replace("void ", ""), tag),
                   sub, "%s_eval_kernel_pruned" % tag, stages=True)
print("written to", out_dir)
