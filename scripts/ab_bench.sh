#!/bin/bash
# usage: scripts/ab_bench.sh "<defs A>" "<defs B>" ...   -- rebuilds with each set of -D knobs and prints kernel times
for defs in "$@"; do
  TIF_NVCC_DEFS="$defs" python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
  TIF_NVCC_DEFS="$defs" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
  python - "$defs" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
print("[%s] pairs/s %.0f  step %.3f ms  " % (sys.argv[1], d["value"], d["ms_per_step"]) + "  ".join("%s %.3f" % (k[:22], v["ms"]) for k, v in d["roofline"]["kernels"].items()))
PY
done
python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
