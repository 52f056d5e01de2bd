#!/bin/bash
# per-kernel time split of the exact filter mode (ncu launch list, cold-cache per-launch times)
for cfg in "4294967296 4096 64 4" "1073741824 256 200 10" "1073741824 1024 64 4"; do
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/dna_l.csv python tools/filter_profile.py $cfg > /dev/null 2>&1
  echo "== $cfg"
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/dna_l.csv")) if len(r)>10][1:]
seen=set()
for r in rows[-16:]:
    key=(r[4],r[12])
    if key in seen: continue
    seen.add(key)
    print(r[4][:50], r[12], r[14])
PY
done
