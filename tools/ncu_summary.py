"""Summarises `ncu --set full` reports into profiles/: one CSV of the headline metrics per kernel, the executed
instruction mix per pixel, and profiles/traffic.json (dram bytes per launch, read by bench.py).

usage: python tools/ncu_summary.py <round-tag> <workload> <pixels-per-launch> <report.ncu-rep> [<report2> ...]"""
import csv
import re
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag, workload, npx = sys.argv[1], sys.argv[2], float(sys.argv[3])
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in sys.argv[4:]:
      hdr, units, launches = raw(rep)
      seen = set()
      for li, vals in enumerate(launches):
        full = vals[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
        kname = full.split("<")[0]
        m = re.search(r"(ring_pointwise_kernel|ring_reduce_kernel|rein_ring_kernel)<(?:sb::)?(\w+(?:<\d>)?)", full)
        if m:                                           # the ring templates: keep the Op in the name
            kname = f"{m.group(1)}<{m.group(2)}>"
        if kname in seen:                               # first profiled launch of every kernel in the report
            continue
        seen.add(kname)
        base = os.path.join(ROOT, "profiles", f"{tag}_{workload}_{kname}".replace("<", "_").replace(">", ""))
        with open(base + "_ncu.csv", "w") as f:
            f.write(f"# ncu --set full --clock-control none, {workload}, first profiled launch; from {os.path.basename(rep)}\n")
            f.write(f"Kernel Name,,{vals[hdr.index('Kernel Name')]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k},{units[i]},{vals[i]}\n")
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                    try:
                        stalls.append((float(vals[i]), h))
                    except ValueError:
                        pass
            for v, h in sorted(stalls, reverse=True)[:8]:
                f.write(f"{h},ratio,{v}\n")
        rd = float(vals[hdr.index("dram__bytes_read.sum")]) * UNIT[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(vals[hdr.index("dram__bytes_write.sum")]) * UNIT[units[hdr.index("dram__bytes_write.sum")]]
        # stamped with the sources it was captured from: bench.py reports it only while those are the sources on disk
        sys.path.insert(0, ROOT)
        from stainlib_b200.build import source_sha16
        traffic.setdefault(workload, {})[kname] = {"dram_bytes": int(rd + wr), "src_sha16": source_sha16(), "report": os.path.basename(rep)}
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{li + 1}"], capture_output=True, text=True).stdout
        mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), str(npx), "30"], input=src, capture_output=True, text=True).stdout
        with open(base + "_opmix.txt", "w") as f:
            f.write(f"# executed thread-instructions per pixel, {workload}, {os.path.basename(rep)}\n" + mix)
        print(kname, "dram bytes/launch", int(rd + wr))
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
