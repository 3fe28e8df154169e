"""Summarise an `ncu --csv` launch list: per-launch duration and DRAM bytes by kernel."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
data = [r for r in rows if r and r[0].isdigit()]
d = defaultdict(dict)
for r in data:
    d[int(r[0])][r[12]] = float(r[14].replace(',', ''))
    d[int(r[0])]['k'] = r[4].split('(')[0][-44:]
    d[int(r[0])]['grid'] = r[8]
for i in sorted(d):
    e = d[i]
    print(i, e['k'], e['grid'], "us=%.1f" % (e.get('gpu__time_duration.sum', 0) / 1e3),
          "rdMB=%.1f wrMB=%.1f" % (e.get('dram__bytes_read.sum', 0) / 1e6, e.get('dram__bytes_write.sum', 0) / 1e6))
