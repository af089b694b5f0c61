"""Per-kernel table from .ncu-rep captures -> JSON (profiles/r2_kernel_table.json, read by bench.py's roofline key).

  python tools/ncu_kernel_table.py OUT.json hot.ncu-rep [sw.ncu-rep SW_CELLS]

Per kernel (the LAST launch of each name in the capture): ncu duration (cold-cache, serialised -- shares, not absolutes),
dram bytes read + written, issue-slot utilisation, active lanes per instruction, resident warps, warp instructions."""
import csv, io, json, subprocess, sys


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1)


def table(rep):
    hdr, units, rows = rows_of(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()

        def g(metric):
            return num(r[ix[metric]]) if metric in ix else None

        dr, dw = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        t = g("gpu__time_duration.sum")
        tu = units[ix["gpu__time_duration.sum"]]
        out[name] = {
            "ncu_us": t * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(tu, 1) if t is not None else None,
            "dram_bytes": (to_bytes(dr, units[ix["dram__bytes_read.sum"]]) + to_bytes(dw, units[ix["dram__bytes_write.sum"]]))
            if dr is not None and dw is not None else None,
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_inst": g("smsp__inst_executed.sum"),
            "thread_inst": g("smsp__thread_inst_executed.sum"),
            "registers": g("launch__registers_per_thread"),
            "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
        }
    return out


if __name__ == "__main__":
    out_path, hot = sys.argv[1], sys.argv[2]
    t = table(hot)
    if len(sys.argv) >= 5:
        sw = table(sys.argv[3])
        cells = float(sys.argv[4])
        for k, v in sw.items():
            if "sw_kernel" in k:
                ti = v.get("thread_inst") or ((v.get("warp_inst") or 0) * (v.get("lanes_per_inst") or 0))
                v["thread_inst_per_cell"] = ti / cells if cells else None
                v["cells"] = cells
                t["sw_kernel"] = v
    json.dump(t, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(t, indent=1, sort_keys=True))
