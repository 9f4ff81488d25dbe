"""Turn ncu output into the text summaries kept under profiles/.

  ncu_summary.py launches <launches.csv> "<title line>"      -> markdown table: kernel | launches | total ms | share | max us
  ncu_summary.py full <report.ncu-rep> "<title line>"        -> the metrics the roofline discussion uses, one block per captured launch

The CSV comes from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> <cmd>`; the report from
`ncu --set full --clock-control none --import-source on -o <report> <cmd>` (both run on the GPU box; this script runs anywhere ncu is installed)."""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("sp::dev::", "")


def launches(path, title):
    rows = [l for l in open(path, errors="replace") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        t = tot[short(r["Kernel Name"])]
        t[0] += 1; t[1] += us; t[2] = max(t[2], us)
    total = sum(t[1] for t in tot.values())
    print("# " + title + "\n")
    print("| kernel | launches | total ms | share | max us |\n|---|---|---|---|---|")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% | %.1f |" % (k, t[0], t[1] / 1e3, 100 * t[1] / total, t[2]))
    print("\ntotal %.3f ms over %d launches" % (total / 1e3, sum(t[0] for t in tot.values())))


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    print("# " + title)
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        print("Kernel Name = " + d.get("Kernel Name", "?"))
        for m in METRICS:
            if m in d:
                print("%s = %s %s" % (m, d[m], units[hdr.index(m)]))
        st = []
        for h in hdr:
            mm = STALL.match(h)
            if mm and "not_issued" not in h and d[h]:
                try:
                    st.append((float(d[h].replace(",", "")), mm.group(1) or mm.group(2)))
                except ValueError:
                    pass
        seen = set()
        for v, n in sorted(st, reverse=True):
            if n in seen:
                continue
            seen.add(n)
            if len(seen) > 7:
                break
            print("stall %.2f %s" % (v, n))
        print()





def traffic_json(path, out_json):
    """`full` captures -> {"<family>": {dram_bytes, us, kernel, what}} of the FIRST captured launch of each kernel family: what bench.py reports as
    roofline.traffic (profiles/r02_ncu_traffic.json)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    res = {}

    def num(d, key):
        v = float(d[key].replace(",", ""))
        u = units[hdr.index(key)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        name = short(d.get("Kernel Name", "?"))
        fam = "sc_fold_eval" if "sc_fold_eval" in name else ("msm_rows" if "msm_rows" in name else ("ipa_msm" if "ipa_msm" in name else name))
        if fam in res:
            continue
        res[fam] = {"kernel": name, "dram_bytes": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"), "us": num(d, "gpu__time_duration.sum"),
                    "grid": d.get("launch__grid_size"), "regs": d.get("launch__registers_per_thread"),
                    "warp_instructions": float(d.get("smsp__inst_executed.sum", "0").replace(",", "") or 0),
                    "issue_active_pct": float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "0").replace(",", "") or 0),
                    "what": "first captured launch of this kernel in a SNARK::prove at 2^20 (ncu --set full --clock-control none)"}
    json.dump(res, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic_json}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
