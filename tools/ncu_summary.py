"""Summarise an `ncu --csv` launch list (one row per launch and metric) into a per-kernel table for ONE forward pass:
time share, DRAM GB/s and % of the measured HBM peak, tensor-pipe %.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 3000 --csv --log-file launches.csv \
        python bench.py --steps 1 --warmup 1
    python tools/ncu_summary.py launches.csv > profiles/rNN_launches_summary.csv

The window is the launches from the 2nd `preprocess_kernel` to the 3rd (one whole forward, warm).  ncu serialises the launches
and replays them with cold caches: compare SHARES, not absolute times, with the CUDA-event numbers of bench.py."""
import csv, json, os, re, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("pf::", "").replace("(int)", "")
    return name


def main(path):
    rows = OrderedDict()
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        d = rows.setdefault(i, {"name": short(r["Kernel Name"])})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
            d[r["Metric Name"] + ".unit"] = r["Metric Unit"]
        except ValueError:
            pass
    ids = sorted(rows)
    pre = [i for i in ids if rows[i]["name"].startswith("preprocess_kernel")]
    lo, hi = (pre[1], pre[2]) if len(pre) >= 3 else (ids[0], ids[-1] + 1)
    win = [rows[i] for i in ids if lo <= i < hi]
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = None

    def ns(d):
        v, u = d.get("gpu__time_duration.sum", 0.0), d.get("gpu__time_duration.sum.unit", "ns")
        return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)

    def byts(d, k):
        v, u = d.get(k, 0.0), d.get(k + ".unit", "byte")
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    agg = OrderedDict()
    for d in win:
        a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "bytes": 0.0, "tensor_w": 0.0})
        a["n"] += 1
        a["ns"] += ns(d)
        a["bytes"] += byts(d, "dram__bytes_read.sum") + byts(d, "dram__bytes_write.sum")
        a["tensor_w"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * ns(d)
    total = sum(a["ns"] for a in agg.values())
    print(f"# one forward: launches {lo}..{hi - 1} of {os.path.basename(path)}; cold-cache, serialised: compare SHARES")
    print(f"# hbm_peak_gbps (MEASURED_PEAKS.json) = {hbm_peak}")
    print("share_pct,total_ms,launches,dram_GB,dram_GBps,dram_pct_of_measured_peak,tensor_pipe_pct_time_weighted,kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        gbps = a["bytes"] / a["ns"] if a["ns"] else 0.0
        print("%.2f,%.3f,%d,%.3f,%.0f,%s,%.1f,%s" % (100 * a["ns"] / total, a["ns"] / 1e6, a["n"], a["bytes"] / 1e9, gbps,
                                                   ("%.1f" % (100 * gbps / hbm_peak)) if hbm_peak else "", a["tensor_w"] / a["ns"] if a["ns"] else 0.0, k))
    print("100.00,%.3f,%d,,,,,TOTAL" % (total / 1e6, sum(a["n"] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
