"""Summarise ncu outputs into profiles/: (1) the per-launch time list of one timed bench step
(shares per kernel), (2) selected raw metrics of an `ncu --set full` report."""
import collections, csv, re, subprocess, sys

def launch_shares(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        tot[name] += v; cnt[name] += 1
        rows.append((r["ID"], name, r.get("Grid Size", ""), v))
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list of one timed bench.py step (cold-cache, serialised: compare shares)\n")
        f.write(f"# total {T/1e6:.3f} ms over {sum(cnt.values())} launches\n")
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{k},{cnt[k]},{v/1e6:.4f},{100*v/T:.2f}\n")
        f.write("\n# first V-cycle of the step, launch by launch\nid,kernel,grid,us\n")
        for i, (id_, name, grid, v) in enumerate(rows[:60]):
            f.write(f"{id_},{name},{grid},{v/1e3:.2f}\n")

def raw_metrics(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(txt.splitlines()))
    hdr = rd[0]
    want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
            "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
            "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_sample_buffers"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(out, "w") as f:
        f.write(",".join(w for w, _ in idx) + "\n")
        f.write(",".join(rd[1][i] for _, i in idx) + "\n")
        for r in rd[2:]:
            f.write(",".join('"%s"' % r[i] if "," in r[i] else r[i] for _, i in idx) + "\n")

def setup_summary(path, out):
    """Per-kernel totals of one AMG setup (ncu launch list with a few throughput metrics): time share,
    launches, and the metrics of each kernel's longest launch."""
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(dict))   # name -> id -> metric -> value
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        per[name][r["ID"]][r["Metric Name"]] = v
    tot = {k: sum(m.get("gpu__time_duration.sum", 0.0) for m in ids.values()) for k, ids in per.items()}
    T = sum(tot.values())
    cols = ["sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum"]
    with open(out, "w") as f:
        f.write("# ncu launch list of ONE AMG setup (7-pt 256^3, one GPU; cold-cache, serialised: compare shares)\n")
        f.write(f"# total {T/1e6:.2f} ms over {sum(len(v) for v in per.values())} launches\n")
        f.write("kernel,launches,total_ms,share_pct,longest_ms," + ",".join(c.split(".")[0] for c in cols) + "\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            big = max(per[k].values(), key=lambda m: m.get("gpu__time_duration.sum", 0.0))
            f.write(f"{k},{len(per[k])},{v/1e6:.3f},{100*v/T:.2f},{big.get('gpu__time_duration.sum', 0)/1e6:.3f}," +
                    ",".join(f"{big.get(c, float('nan')):.6g}" for c in cols) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launch_shares(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "setup":
        setup_summary(sys.argv[2], sys.argv[3])
    else:
        raw_metrics(sys.argv[2], sys.argv[3])
