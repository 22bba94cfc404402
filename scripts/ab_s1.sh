#!/bin/bash
# usage: scripts/ab_s1.sh "<defs A>" "<defs B>" ...  -- stage-1 kernel time at 2K and 4K for each set of -D knobs
for defs in "$@"; do
  TIF_NVCC_DEFS="$defs" python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
  for hb in "1024 32" "1920 16" "640 32"; do
    set -- $hb
    TIF_NVCC_DEFS="$defs" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --height $1 --batch $2 > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
    python - "$defs" $1 <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
k = d["roofline"]["kernels"]["k_erp2ico"]
print("[%s] H=%s k_erp2ico %.3f ms (%.1f %% of HBM)" % (sys.argv[1], sys.argv[2], k["ms"], 100 * k["frac"]))
PY
  done
done
python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
