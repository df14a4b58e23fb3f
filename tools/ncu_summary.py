"""Condense an .ncu-rep (ncu --set full) into one line of roofline-relevant metrics per launch.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tc_wavefronts_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%")]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: per-launch metrics (ncu --set full --clock-control none; cold cache, serialised)")
for r in rows[2:]:
    parts = []
    for key, label in want:
        i = idx.get(key)
        if i is None:
            continue
        v = r[i]
        if key == "Kernel Name":
            v = v.split("(")[0].replace("void ", "")
        parts.append(f"{label}={v}{(' ' + units[i]) if units[i] and key != 'Kernel Name' else ''}")
    print(" | ".join(parts))
