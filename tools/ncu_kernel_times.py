"""Print the last N kernel launches (name, microseconds, DRAM read / write MB) of an ncu --csv launch list."""
import csv, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault(int(r[ii]), {"k": r[ki][:58]})[r[mi]] = float(r[vi].replace(",", ""))
for i in sorted(d)[-n:]:
    e = d[i]
    print(f"{i:4d} {e['k']:58s} {e.get('gpu__time_duration.sum', 0) / 1e3:9.1f} us  "
          f"rd {e.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  wr {e.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB")
