"""Summarise `ncu --set full` reports (tools/ncu_full_kernels.sh) into one JSON: the metrics the design discussion cites.
    python tools/ncu_full_summary.py gpurun_out/r02_full_*.ncu-rep > profiles/r02_ncu_full_kernels.json
Reads each report with `ncu -i <rep> --page raw --csv` (works without a GPU)."""
import csv, io, json, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main(paths):
    out = {}
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        e = {"kernel": d["Kernel Name"][0][:160], "grid": d.get("Grid Size", ("", ""))[0], "block": d.get("Block Size", ("", ""))[0]}
        for k in KEYS:
            if k in d and d[k][0] != "":
                e[k] = ("%s %s" % d[k]).strip()
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            e["dram_bytes_per_launch"] = int(to_bytes(*rd) + to_bytes(*wr))
        out[os.path.basename(p).replace(".ncu-rep", "")] = e
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])
