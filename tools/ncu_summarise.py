"""Turns a `ncu --set full` report (gpurun_out/<tag>_top.ncu-rep) into the small JSON summary committed under profiles/.

    python tools/ncu_summarise.py gpurun_out/r02_top.ncu-rep profiles/r02_bw3_kernel_ncu_full_summary.json "<source cmd>"

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and keeps the metrics the roofline discussion uses.
`dram_bytes_per_launch` (read + write) is what bench.py reports as `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(val: str, unit: str) -> float:
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    return float(val.replace(",", "")) * mul


def main():
    rep, out, src = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        rec = dict(zip(header, r))
        k = {"Kernel Name": [rec.get("Kernel Name", ""), ""]}
        for name in KEEP:
            if name in rec:
                k[name] = [rec[name], units[header.index(name)]]
        rd = to_bytes(*k["dram__bytes_read.sum"]) if "dram__bytes_read.sum" in k else 0.0
        wr = to_bytes(*k["dram__bytes_write.sum"]) if "dram__bytes_write.sum" in k else 0.0
        k["dram_bytes"] = rd + wr
        kernels.append(k)
    summary = {"source": src, "kernels": kernels,
               "dram_bytes_per_launch": sum(k["dram_bytes"] for k in kernels) / max(1, len(kernels))}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in kernels[0].items() if k != "Kernel Name"}, indent=1))


if __name__ == "__main__":
    main()
