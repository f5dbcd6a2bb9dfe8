"""Pick the metrics bench.py / profiles/README.md quote out of an ncu report.
usage: ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > /tmp/raw.csv; python tools/ncu_metrics.py /tmp/raw.csv > profiles/<name>.json
(first profiled launch of the report)"""
import csv, json, sys
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_tensor", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active")
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units, first = rows[0], rows[1], rows[2]
out = {"kernel": first[hdr.index("Kernel Name")]}
for i, h in enumerate(hdr):
    if any(k in h for k in KEEP):
        out[h] = {"unit": units[i], "value": first[i]}
print(json.dumps(out, indent=1))
