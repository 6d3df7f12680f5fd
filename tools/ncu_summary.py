"""Summarise an .ncu-rep (read here on the CPU box): key metrics + hottest SASS lines with their stall reasons.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_hot]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem"]
print(f"## {rep}\n")
for h, u, v in zip(hdr, units, vals):
    if h in want or any(h == w for w in want):
        print(f"- `{h}` = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
repeat = [i for i, r in enumerate(rows) if r == h]   # one table per captured kernel: summarise the first
end = repeat[1] - 1 if len(repeat) > 1 else len(rows)
data = [r for r in rows[2:end] if len(r) == len(h)]
si, so = h.index("# Samples"), h.index("Source")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[si]) for r in data) or 1
print(f"\nhottest SASS lines ({tot} samples):\n")
for r in sorted(data, key=lambda r: -int(r[si]))[:n_hot]:
    reasons = sorted(((int(r[i]), h[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    print(f"- {100*int(r[si])/tot:5.1f}%  `{r[so].strip()[:64]}`  {', '.join(f'{n} {c}' for c, n in reasons)}")
