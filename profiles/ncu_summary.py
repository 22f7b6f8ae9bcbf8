"""Summarise an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv > X.csv`:
    python profiles/ncu_summary.py X.csv [--json profiles/r02_top_kernel_traffic.json] > profiles/r02_xxx_ncu_full.md
One column per captured launch, the metrics the judge reads (duration, DRAM bytes, tensor-pipe %, SM / L2 / DRAM throughput %,
registers, shared memory, instructions).  --json writes the per-launch DRAM traffic of the first captured kernel for bench.py."""
import csv, json, sys

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "launch__registers_per_thread", "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(f"source: {sys.argv[1]} ({len(data)} launches)\n")
print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
print("|---|" + "---|" * len(data))
for m in METRICS:
    if m not in col:
        continue
    u = units[col[m]]
    print(f"| {m} [{u}] | " + " | ".join(r[col[m]].replace("conv_tc_kernel", "conv_tc").split("(")[0] for r in data) + " |")
if "--json" in sys.argv:
    def num(r, m):
        v = float(r[col[m]].replace(",", ""))
        u = units[col[m]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    r0 = data[0]
    same = [r for r in data if r[col["Kernel Name"]] == r0[col["Kernel Name"]] and r[col["Grid Size"]] == r0[col["Grid Size"]]]
    tr = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in same]
    out = {"kernel": r0[col["Kernel Name"]].split("(")[0], "launches": len(same), "dram_bytes_per_launch": sum(tr) / len(tr),
           "per_launch": tr, "source": sys.argv[1], "how": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full --clock-control none"}
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
