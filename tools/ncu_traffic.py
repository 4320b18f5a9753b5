"""profiles/<tag>_traffic.json + raw-metric CSV rows from `ncu --set full` reports of the bench kernels (development aid).

    python tools/ncu_traffic.py <tag> report1.ncu-rep [report2.ncu-rep ...]

For every kernel in the reports: DRAM bytes read + written per launch (-> the `traffic` field bench.py puts in its roofline
objects, keyed by the C-ABI call the kernel belongs to) and a CSV of the headline raw metrics (profiles/<tag>_ncu_raw.csv)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, reps = sys.argv[1], sys.argv[2:]
CALL = {"gconv_stack_pair": "eqb_gconv_stack_run", "gconv_stack_tc": "eqb_gconv_stack_run", "resample_tma_kernel<3, 0>": None,
        "crop_resize": "eqb_crop_resize_aa"}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
traffic, out_rows, warp_seen = {}, [["kernel", "metric", "unit", "value"]], 0
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ik]
        vals = {h: (u, v) for h, u, v in zip(hdr, units, r)}
        for m in KEEP:
            if m in vals:
                out_rows.append([name[:70], m, vals[m][0], vals[m][1]])
        def mb(key):
            u, v = vals[key]
            f = float(v.replace(",", ""))
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tot = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        key = None
        for frag, call in CALL.items():
            if frag in name:
                key = call
        if "resample_tma_kernel" in name or "resample_tma_persistent_kernel" in name:
            key = "eqb_warp_canonicalize" if warp_seen == 0 else "eqb_warp_invert"
            warp_seen += 1
        if key and key not in traffic:
            traffic[key] = tot
json.dump(traffic, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
csv.writer(open(os.path.join(ROOT, "profiles", tag + "_ncu_raw.csv"), "w")).writerows(out_rows)
print(traffic)
