"""Top CUDA source lines of one kernel by warp-stall samples, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > file.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[2]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[3:]:
    if len(r) > ie and r[0] not in ("", "Line No"):
        try:
            data.append((int(r[si]), int(r[ie]), r[0], r[1][:120]))
        except ValueError:
            pass
tot, toti = sum(d[0] for d in data), sum(d[1] for d in data)
print("total samples", tot, "warp instructions", toti)
for d in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}%  inst {d[1]:8d} {100 * d[1] / toti:5.1f}%  L{d[2]}: {d[3]}")
