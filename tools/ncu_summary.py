"""Summarise gpurun_out/*.ncu-rep + the launch-list CSV into profiles/ (tracked).  Run where ncu is installed (no GPU needed)."""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h = r[0]
    rows = []
    for row in r[2:]:
        d = {"kernel": re.sub(r"\(.*", "", row[h.index("Kernel Name")])}
        for k in KEYS:
            if k in h:
                d[k] = f"{row[h.index(k)]} {r[1][h.index(k)]}".strip()
        rows.append(d)
    return rows


def stalls(rep, top=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    isamp, isrc, iexe = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [(int(r[isamp] or 0), r[isrc].strip(), int(r[iexe] or 0)) for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
    tot = max(1, sum(d[0] for d in data))
    g = collections.Counter()
    for s, src, e in data:
        t = src.split()
        g[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += s
    sass = collections.Counter()
    for s, src, e in data:
        for m in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "UBLKCP"):
            if m in src:
                sass[m] += e
    return {"samples": tot, "by_opcode_pct": {k: round(100.0 * v / tot, 1) for k, v in g.most_common(top)}, "sass_executed": dict(sass)}


def launches(csv_path):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u in ("us", "usecond") else v
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        k = re.sub(r"void |unnamed>::|<unnamed>::|q2w::", "", k)
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    return {k: {"launches": n, "total_ms": round(t, 3), "share": round(t / tot, 4), "avg_us": round(1e3 * t / n, 2)} for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])}


if __name__ == "__main__":
    tag = sys.argv[1]
    out = {"note": "ncu --set full --clock-control none captures (cold cache, serialised); shares, not absolutes, are comparable with bench.py's live CUDA-event numbers"}
    for name in sys.argv[3:]:
        rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
        out[name] = {"metrics": raw(rep), "stall_sampling": stalls(rep)}
    lp = os.path.join(ROOT, "gpurun_out", sys.argv[2])
    if os.path.exists(lp):
        out["launch_list"] = {"command": "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --windows 16  (first 1300 launches)", "kernels": launches(lp)}
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json"), "w"), indent=1)
    print(json.dumps(out.get("launch_list", {}), indent=1)[:3000])
    for name in sys.argv[3:]:
        for m in out[name]["metrics"]:
            print(name, {k.split(".")[0][-40:]: v for k, v in m.items() if k in ("kernel", KEYS[0], KEYS[5], KEYS[6], KEYS[8], KEYS[9], KEYS[10])})
        print("   stalls:", out[name]["stall_sampling"]["by_opcode_pct"], out[name]["stall_sampling"]["sass_executed"])
