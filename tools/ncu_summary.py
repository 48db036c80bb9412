"""Writes a text summary (key raw metrics + hottest source lines) of an .ncu-rep into profiles/."""
import csv, io, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none --import-source on)", ""]
for wname in want:
    if wname in hdr:
        i = hdr.index(wname)
        unit = rows[1][i] if len(rows) > 1 else ""
        lines.append(f"{wname} [{unit}]: " + " | ".join(r[i] for r in rows[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
agg = []
for r in srows[3:]:
    if not r or not r[0].strip().isdigit():
        continue
    for i in range(1, len(r) - 3):
        if r[i] == '-' and r[i + 1] == '-' and r[i + 2].isdigit():
            agg.append((int(r[i + 2]), int(r[0]), ','.join(r[1:i]).strip()[:110]))
            break
tot = sum(a[0] for a in agg) or 1
lines += ["", f"hottest CUDA source lines by warp-stall samples (total {tot}; backend_cuda.cu line numbers at capture time):"]
for s, l, text in sorted(agg, reverse=True)[:16]:
    lines.append(f"  {100*s/tot:5.1f}%  L{l:<5d} {text}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
