#!/usr/bin/env python
"""Turns the ncu captures a GPU round brought back (gpurun_out/<tag>/) into the small text / CSV
summaries that are committed under profiles/.  Usage: python tools/summarize_profiles.py <tag> [round]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize(rep, out_name, title):
    if not os.path.exists(rep):
        return None
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s" % title, "# source: ncu --set full --clock-control none --import-source on (replayed, serialised, cold cache):",
             "# durations are for shares only, never bench values", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        lines.append("kernel: " + name)
        for k in KEYS:
            if k in idx:
                lines.append("  %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_write.sum"]]]
            traffic[name.split("(")[0]] = rd + wr
            lines.append("  dram traffic per launch (read + write)                                    %.0f bytes" % (rd + wr))
        except Exception:
            pass
        lines.append("")
    # stall reasons / instruction mix per kernel from the source page
    src_rows = ncu_csv(rep, "source")
    if len(src_rows) > 2:
        hdr2 = src_rows[1]
        if "Warp Stall Sampling (All Samples)" in hdr2:
            i_s, i_ex, i_src = hdr2.index("Warp Stall Sampling (All Samples)"), hdr2.index("Instructions Executed"), hdr2.index("Source")
            cols = [(i, h) for i, h in enumerate(hdr2) if h.startswith("stall_") and "Not Issued" not in h]
            data = [r for r in src_rows[2:] if len(r) > i_s and r[i_s].isdigit()]
            tot = sum(int(r[i_s]) for r in data) or 1
            agg = collections.Counter()
            ops = collections.Counter()
            for r in data:
                for i, h in cols:
                    if r[i].isdigit():
                        agg[h[6:]] += int(r[i])
                t = r[i_src].strip().split()
                if t:
                    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                    ops[op.split(".")[0]] += int(r[i_ex])
            lines.append("warp stall sampling, all captured kernels (share of samples):")
            lines += ["  %-20s %5.1f%%" % (k, 100.0 * v / tot) for k, v in agg.most_common(10)]
            tex = sum(ops.values()) or 1
            lines.append("executed instruction mix (share of warp instructions):")
            lines += ["  %-10s %5.1f%%" % (k, 100.0 * v / tex) for k, v in ops.most_common(14)]
    open(os.path.join(dst, out_name), "w").write("\n".join(lines) + "\n")
    return traffic


commit = None
if os.path.exists(os.path.join(src, "commit.txt")):
    commit = open(os.path.join(src, "commit.txt")).read().strip()
t1 = summarize(os.path.join(src, "prof_fused4.ncu-rep" if os.path.exists(os.path.join(src, "prof_fused4.ncu-rep")) else "prof_fused.ncu-rep"),
               "%s_ncu_fused_kernel.txt" % rnd, "music4_fused_kernel (default path for M=4, n=1): BASELINE config 2, 10 000 windows per launch")
t2 = summarize(os.path.join(src, "prof.ncu-rep"), "%s_ncu_unfused_kernels.txt" % rnd,
               "unfused three-kernel path (MUSIC_B200_FUSED=0): cov4_tma_kernel, eig_kernel, scan_peak1_kernel")
t8 = summarize(os.path.join(src, "prof_fused8.ncu-rep"), "%s_ncu_fused8_kernel.txt" % rnd,
               "music8_fused_kernel (default path for M=8, n=1): BASELINE config 4 shape, 8192 windows per launch")
t4 = summarize(os.path.join(src, "prof_covn_c4.ncu-rep"), "%s_ncu_covN_M8.txt" % rnd,
               "covN_tma_kernel<8> (K1 for M = 8): BASELINE config 4 shape, 8192 windows per launch")
t5 = summarize(os.path.join(src, "prof_c5.ncu-rep" if os.path.exists(os.path.join(src, "prof_c5.ncu-rep")) else "prof_covn_c5.ncu-rep"),
               "%s_ncu_config5_kernels.txt" % rnd, "covN_tma_kernel<16> and eig_coop_kernel<16> (M = 16): BASELINE config 5 shape")
for f, name in (("bench_c3.json", "%s_bench_config3.json" % rnd), ("bench_c4.json", "%s_bench_config4.json" % rnd),
                ("bench_c5.json", "%s_bench_config5.json" % rnd), ("launches.csv", "%s_launches.csv" % rnd), ("microbench.log", "%s_microbench.txt" % rnd),
                ("bench.json", "%s_bench_1gpu.json" % rnd), ("bench_ref.json", "%s_bench_reference_arm.json" % rnd),
                ("bench_MUSIC_B200_FUSED_0.json", "%s_bench_unfused.json" % rnd), ("pytest_gpu.log", "%s_pytest_gpu.txt" % rnd),
                ("gpu.txt", "%s_gpu.txt" % rnd), ("fused_trace.txt", "%s_fused_trace.txt" % rnd), ("smoke.log", "%s_smoke.txt" % rnd)):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), os.path.join(dst, name))
if t1:
    tr = {"config2": next(iter(t1.values())), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
          "kernel": next(iter(t1.keys())), "capture": "%s_ncu_fused_kernel.txt" % rnd, "capture_commit": commit}
    if t8:
        tr["config4"] = next(iter(t8.values()))
        tr["config4_kernel"] = next(iter(t8.keys()))
    if t5:
        tr["config5_kernels"] = t5
    json.dump(tr, open(os.path.join(dst, "roofline_traffic.json"), "w"), indent=1)
print(os.listdir(dst))
