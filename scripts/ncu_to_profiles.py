"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the artefacts kept under profiles/: the full raw-metric CSV
and a small JSON (duration, DRAM bytes per launch, pipe utilisation) that bench.py attaches to its roofline line.
Usage: python scripts/ncu_to_profiles.py <rep> <out_stem> [source-note]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "launch__registers_per_thread": "registers",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def main(rep, stem, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(stem + ".csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    out = {"source": stem.split("/")[-1] + ".csv", "note": note}
    for i, name in enumerate(h):
        if name == "Kernel Name":
            out["kernel"] = v[i]
        if name in KEYS:
            out[KEYS[name]] = float(v[i].replace(",", "")) * UNIT.get(u[i], 1.0)
    out["dram_bytes_per_launch"] = out.get("dram_read", 0) + out.get("dram_write", 0)
    json.dump(out, open(stem + ".json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
