"""Turn an .ncu-rep into the small CSV / JSON summaries committed under profiles/.

    python scripts/summarize_ncu.py gpurun_out/x.ncu-rep profiles/r1_x   (needs `ncu` on PATH, no GPU)
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def to_bytes(val, unit):
    v = float(val)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in ix]
    with open(out + "_kernels.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[ix[c]] for c in cols])
        for r in data:
            w.writerow([r[ix[c]] for c in cols])
    traffic = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        traffic.setdefault(name, []).append(rd + wr)
    with open(out + "_dram_traffic.json", "w") as f:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, f, indent=1)
    print(f"wrote {out}_kernels.csv / _dram_traffic.json ({len(data)} launches)")


if __name__ == "__main__":
    main()
