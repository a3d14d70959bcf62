"""Summary of an `ncu --set full` capture: per kernel the duration, DRAM traffic, issue utilisation, occupancy and top stall
reasons; with --traffic-json also DRAM bytes per record (dram__bytes_read.sum + dram__bytes_write.sum over the records of the
captured run), the figure bench.py multiplies by the records of its own run for `roofline.traffic`.

    python tests/tools/ncu_summary.py gpurun_out/r2_prof_20M.ncu-rep --records 38264724 --traffic-json profiles/r2_traffic.json > profiles/r2_ncu_summary.txt
"""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

SHORT = {"k_classify_tiles": "k_classify", "k_edges_generic": "k_edges_generic", "k_cov_compact": "k_cov_compact", "k_cov_count_tiles": "k_cov_count",
         "k_seed_islands": "k_seed_islands", "k_wire_decode": "k_wire_decode"}


def short_name(name: str) -> str:
    for k, v in SHORT.items():
        if k in name:
            return v
    if "k_assign_tiles" in name:
        m = re.search(r"k_assign_tiles<\(bool\)(\d), \(bool\)(\d)", name) or re.search(r"k_assign_tiles<(\w+), (\w+)", name)
        if m:
            d, e = m.group(1) in ("1", "true"), m.group(2) in ("1", "true")
            return "k_assign_depth" if d and not e else ("k_assign_edges" if e and not d else "k_assign")
        return "k_assign"
    return name.split("(")[0][-40:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--records", type=int, default=0)
    ap.add_argument("--traffic-json")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, default=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return default
    units = rows[1]
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") or h.startswith("smsp__average_warps_issue_stalled_")]
    if not stall_cols:
        stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    traffic = {}
    print("# %s%s" % (a.rep, (" (%d records)" % a.records) if a.records else ""))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        dur = get(r, "gpu__time_duration.sum")
        du = units[col["gpu__time_duration.sum"]]
        ms = dur / 1e6 if du in ("ns", "nsecond") else (dur / 1e3 if du in ("us", "usecond") else dur)
        rd, wr = get(r, "dram__bytes_read.sum", 0.0), get(r, "dram__bytes_write.sum", 0.0)
        for nm, v in (("dram__bytes_read.sum", rd), ("dram__bytes_write.sum", wr)):
            pass
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale.get(units[col["dram__bytes_read.sum"]], 1); wr *= scale.get(units[col["dram__bytes_write.sum"]], 1)
        stalls = sorted(((get(r, c, 0.0), c) for c in stall_cols), reverse=True)[:4]
        sn = short_name(name)
        print("%-18s %8.3f ms  dram %7.3f GB (r %.3f w %.3f)%s  %.0f GB/s  issue %.1f%%  warps_active %.1f%%  regs %d  grid %s  top stalls: %s" % (
            sn, ms, (rd + wr) / 1e9, rd / 1e9, wr / 1e9, ("  %.1f B/record" % ((rd + wr) / a.records)) if a.records else "", (rd + wr) / 1e9 / (ms / 1e3),
            get(r, "sm__inst_issued.avg.pct_of_peak_sustained_active", get(r, "smsp__issue_active.avg.pct", 0.0)) or 0.0,
            get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) or 0.0, int(get(r, "launch__registers_per_thread", 0) or 0), r[col["Grid Size"]] if "Grid Size" in col else "?",
            ", ".join("%s %.2f" % (re.sub(r".*issue_stalled_|_per_warp_active.pct|\.ratio|\.pct", "", c), v) for v, c in stalls)))
        if a.records and sn not in traffic:
            traffic[sn] = (rd + wr) / a.records
    if a.traffic_json:
        json.dump({"source": a.rep.split("/")[-1], "records": a.records, "dram_bytes_per_record": traffic}, open(a.traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()
