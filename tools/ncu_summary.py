"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full --clock-control none`) into the small
JSON files committed under profiles/: per launch the metrics the roofline discussion needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "source / note text" > profiles/<name>.json
Runs where ncu is installed (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
]


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    out = {"source": rep, "note": note, "launches": []}
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]] if "Kernel Name" in col else "?"}
        for k in KEEP:
            if k in col and r[col[k]] != "":
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    v = r[col[k]]
                rec[k] = {"value": v, "unit": units[col[k]]}
        out["launches"].append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
