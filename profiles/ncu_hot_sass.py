"""Top SASS instructions by stall samples from an `ncu --page source --csv` dump; prints neighbours for context."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = rows[2:]
total = sum(int(r[isamp]) for r in body)
top = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
print("total samples", total)
for i in sorted(top):
    r = body[i]
    print(f"{i:5d} {100 * int(r[isamp]) / total:5.1f}%  exec={int(r[iexec]):>10d}  {r[isrc].strip()[:110]}")
