"""Condenses an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
of `bench.py --steps S --warmup W ...` into per-kernel shares and the DRAM traffic per step of each stage:

    python tools/launch_list_summary.py gpurun_out/launches.csv STEPS_TOTAL [out.json] > profiles/rNN_launch_summary.txt

STEPS_TOTAL = warm-up + timed + end-to-end passes of the profiled command (every pass launches the same kernels).
"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
steps = float(sys.argv[2])
per = collections.defaultdict(lambda: {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
launches = {}
one = collections.defaultdict(lambda: {"ns": 0.0, "bytes": 0.0})   # per launch id
for r in rows:
    lid, name, metric, unit, val = r[0], r[4], r[12], r[13], float(r[14].replace(",", ""))
    short = re.sub(r"<unnamed>::|void |\(.*", "", name)
    short = re.sub(r"<.*", "", short)
    d = per[short]
    if lid not in launches:
        launches[lid] = short
        d["n"] += 1
    if metric == "gpu__time_duration.sum":
        ns = val * (1e3 if unit == "usecond" else 1e6 if unit == "msecond" else 1.0)
        d["ns"] += ns
        one[lid]["ns"] += ns
    elif metric == "dram__bytes_read.sum":
        b = val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["rd"] += b
        one[lid]["bytes"] += b
    elif metric == "dram__bytes_write.sum":
        b = val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["wr"] += b
        one[lid]["bytes"] += b
tot = sum(d["ns"] for d in per.values())
print(f"{len(launches)} launches, {tot / 1e6:.2f} ms of kernel time over {steps:g} passes = {tot / 1e6 / steps:.2f} ms per pass "
      f"(ncu serialises the launches and runs them cold: compare shares, not absolutes)")
print(f"{'kernel':28s} {'launches/pass':>13s} {'ms/pass':>9s} {'share':>7s} {'DRAM rd MB/pass':>16s} {'wr MB/pass':>11s}")
for k, d in sorted(per.items(), key=lambda x: -x[1]["ns"]):
    print(f"{k:28s} {d['n'] / steps:13.1f} {d['ns'] / 1e6 / steps:9.3f} {100 * d['ns'] / tot:6.1f}% {d['rd'] / 1e6 / steps:16.1f} {d['wr'] / 1e6 / steps:11.1f}")
SNN = ("k_hist", "k_scan", "k_scatter", "k_sort", "k_emit", "k_row_work", "k_snn", "k_low", "k_row_total", "k_max_indeg",
       "k_info", "k_i64")
stage = {"snn_stage": 0.0, "k_knn_tc": 0.0, "all": 0.0}
for k, d in per.items():
    b = (d["rd"] + d["wr"]) / steps
    stage["all"] += b
    if k.startswith(SNN):
        stage["snn_stage"] += b
    if k.startswith("k_knn_tc"):
        stage["k_knn_tc"] += b
# the pass-2 launch of the grouped search = the longest k_knn_tc launch of every pass (bench.py's roofline kernel)
tc = sorted((v for lid, v in one.items() if launches[lid].startswith("k_knn_tc")), key=lambda v: -v["ns"])[:int(steps)]
if tc:
    stage["k_knn_tc_pass2_launch"] = sum(v["bytes"] for v in tc) / len(tc)
    print(f"k_knn_tc pass-2 launch (longest of each pass): {sum(v['ns'] for v in tc) / len(tc) / 1e6:.3f} ms, "
          f"{stage['k_knn_tc_pass2_launch'] / 1e6:.0f} MB of DRAM traffic per launch")
print("DRAM bytes per pass:", {k: f"{v / 1e6:.0f} MB" for k, v in stage.items()})
if len(sys.argv) > 3:
    json.dump({"snn_stage": stage["snn_stage"], "k_knn_tc": stage["k_knn_tc"], "all_kernels": stage["all"],
               "k_knn_tc_pass2_launch": stage.get("k_knn_tc_pass2_launch"),
               "source": "ncu launch list of `bench.py --steps 2 --warmup 1 --e2e-steps 1` on one B200 (summed over every launch "
                         "of a pass: dram__bytes_read.sum + dram__bytes_write.sum)"}, open(sys.argv[3], "w"), indent=1)
