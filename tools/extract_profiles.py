#!/usr/bin/env python
"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:
   profiles/<tag>_launches.csv        every launch with its device time (cold-cache, serialised: compare shares)
   profiles/<tag>_launch_summary.txt  per-kernel mean/min and share of the step
   profiles/<tag>_full_<kernel>.csv   selected raw metrics of the --set full capture
   profiles/roofline_traffic.json     dram bytes (read + write) per launch, read by bench.py
usage: python tools/extract_profiles.py r01"""
import collections, csv, json, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)

# ---- launch list
src = os.path.join(go, "launches_%s.csv" % tag)
lines = open(src).read().splitlines()
i0 = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
open(os.path.join(pr, "%s_launches.csv" % tag), "w").write("\n".join(lines[i0:]) + "\n")
rows = list(csv.DictReader(lines[i0:]))
d = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0] + " grid" + r["Grid Size"] + " block" + r["Block Size"]
    d.setdefault(k, []).append(float(r["Metric Value"].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in d.values())
with open(os.path.join(pr, "%s_launch_summary.txt" % tag), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none ; python bench.py --steps 5 --warmup 3 --no-cpu-baseline\n")
    f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n")
    f.write("%-70s %5s %9s %9s %7s\n" % ("kernel", "n", "mean_us", "min_us", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-70s %5d %9.1f %9.1f %6.1f%%\n" % (k, len(v), sum(v) / len(v), min(v), 100 * sum(v) / tot))
print(open(os.path.join(pr, "%s_launch_summary.txt" % tag)).read())

want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
traffic = {}
try:
    traffic = json.load(open(os.path.join(pr, "roofline_traffic.json")))
except Exception:
    pass
onchip = {}          # per kernel and launch: shared-memory wavefronts, L2 bytes (bench.py sets them against tools/peaks_lab.cu's peaks)
try:
    onchip = json.load(open(os.path.join(pr, "roofline_onchip.json")))
except Exception:
    pass


# ---- full-set captures (score path, and the UpdateClassifier kernels when captured)
def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        return
    hdr = rr[0]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    units = rr[1]
    seen = set()
    for vals in rr[2:]:
        name = vals[hdr.index("Kernel Name")].split("(")[0]
        # the launch names bench.py reports: no "void", no template arguments; k_pf_update_fast<CH> is launched as k_pf_update
        name = name.replace("void ", "").split("<")[0].replace("k_pf_update_fast", "k_pf_update")
        if name in seen:
            continue
        seen.add(name)
        with open(os.path.join(pr, "%s_full_%s.csv" % (tag, name)), "w") as f:
            f.write("metric,unit,value\n")
            for w, i in idx:
                f.write("%s,%s,%s\n" % (w, units[i], vals[i]))
        def num(m):
            v = float(vals[hdr.index(m)].replace(",", "")); u = units[hdr.index(m)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic[name] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        onchip[name] = {"smem_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                        "smem_bank_conflict_wavefronts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                        "l2_bytes": num("l1tex__m_xbar2l1tex_read_bytes.sum") if "l1tex__m_xbar2l1tex_read_bytes.sum" in hdr else None}


for rep in ("prof_%s.ncu-rep" % tag, "prof_%s_upd.ncu-rep" % tag):
    if os.path.exists(os.path.join(go, rep)):
        full(os.path.join(go, rep))
json.dump(traffic, open(os.path.join(pr, "roofline_traffic.json"), "w"), indent=1)
print("roofline_traffic.json:", traffic)
json.dump(onchip, open(os.path.join(pr, "roofline_onchip.json"), "w"), indent=1)
print("roofline_onchip.json:", onchip)
