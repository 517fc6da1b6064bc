"""Turn `ncu --set full` reports (gpurun_out/full_<kernel>.ncu-rep) into the text summaries committed under profiles/
and the per-launch DRAM traffic table bench.py reports as roofline.traffic.   usage: python scripts/ncu_summary.py r01c"""
import csv, json, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KERNELS = {"rnn_fwd": "rnn_fwd_kernel", "rnn_bwd": "rnn_bwd_kernel", "dw_mm16": "dw_mm16_kernel", "xproj": "bulk_dp_kernel",
           "fft_features": "fft_features_kernel"}
COMMIT = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for short, kern in KERNELS.items():
    rep = f"gpurun_out/full_{kern}.ncu-rep"
    import os
    if not os.path.exists(rep):
        continue
    fft = short == "fft_features"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [f"# ncu --set full ({tag}, commit {COMMIT}): {kern}, " + ("BASELINE config 2 batch (B=512, T=60: 583 680 windows), x + raw outputs" if fft else
           "first launch after warm-up = encoder layer 0 at BASELINE config 2 (B=512, T=60)"),
           f"# command: ncu --set full --clock-control none --import-source on -k regex:{kern} -s 4 -c 1 " +
           ("python scripts/run_fft.py" if fft else "python bench.py --config 2 --steps 2 --no-extra --no-cpu-baseline"), ""]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"{w:90s} {units[i]:16s} {vals[i]}")
    rd = float(vals[hdr.index("dram__bytes_read.sum")]) * UNIT[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(vals[hdr.index("dram__bytes_write.sum")]) * UNIT[units[hdr.index("dram__bytes_write.sum")]]
    traffic[short] = rd + wr
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    if tp in hdr:
        traffic.setdefault("tensor_pipe_active_pct", {})[short] = float(vals[hdr.index(tp)])
    ti = hdr.index("gpu__time_duration.sum")
    traffic.setdefault("time_us", {})[short] = float(vals[ti]) * {"s": 1e6, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(units[ti], 1.0)
    st = [(float(vals[i]), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and vals[i]]
    out += ["", "warp stall reasons (warps stalled per issue-active cycle):"]
    for v, h in sorted(st, reverse=True)[:8]:
        out.append(f"    {v:5.2f} {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
    open(f"profiles/ncu_{short}_{tag}.txt", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(f"profiles/traffic_{tag}.json", "w"), indent=1)
print(traffic)
