"""Write profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) from an
`ncu -i X.ncu-rep --page raw --csv` export of a `--set full` capture, and a compact per-kernel CSV.
  python scripts/ncu_traffic.py raw.csv out_summary.csv [source label]"""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def main():
    raw, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(raw)
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    keep = [k for k in KEEP if k in col]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + keep)
        w.writerow([""] + [units[col[k]] for k in keep])
        for r in body:
            w.writerow([r[col["Kernel Name"]]] + [r[col[k]] for k in keep])

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)

    agg = collections.defaultdict(list)
    for r in body:
        name = r[col["Kernel Name"]]
        grid = r[col["launch__grid_size"]] if "launch__grid_size" in col else ""
        agg[name].append((val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"),
                          float(r[col["sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed"]].replace(",", "")), grid))
    summary = {}
    for name, v in agg.items():
        n = len(v)
        summary[name] = {"launches": n, "avg_us": sum(x[0] for x in v) / n, "dram_read_bytes": sum(x[1] for x in v) / n,
                         "dram_write_bytes": sum(x[2] for x in v) / n, "tensor_pipe_active_pct": sum(x[3] for x in v) / n}
    js = {"source": label, "note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of each kernel in {label} (ncu --set full, B=65,536 rows)",
          "kernels": summary}
    for name, s in summary.items():
        tot = s["dram_read_bytes"] + s["dram_write_bytes"]
        if "gemm2_kernel<0, 1, 3>" in name:      # EPI_PACK, resident A, GELU = mlp.0
            js["mlp0_dram_bytes_per_launch"] = tot
        if "gemm2_kernel<2, 1, 0>" in name:      # EPI_MODLN = adaLN
            js["adaln_dram_bytes_per_launch"] = tot
        if "gemm2_kernel<1, 0, 0>" in name:      # EPI_F32, streamed A = mlp.2
            js["mlp2_dram_bytes_per_launch"] = tot
        if "gemm2_kernel<1, 1, 0>" in name:      # EPI_F32, resident A = folded attention / latent_proj
            js["attn_dram_bytes_per_launch"] = tot
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(js, f, indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    main()
