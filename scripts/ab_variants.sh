#!/bin/bash
# usage (GPU box): scripts/ab_variants.sh [bench flags --] name1 name2 ...  -- runs the bench once with the regular library and once
# per prebuilt variant (360opticalflow-tangentimages_b200/build/variants/<name>.so, see _build.build_variant) and prints the kernel times
FLAGS="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs $ABV_FLAGS"
run() {
  python bench.py $FLAGS > /tmp/abv.json 2>/tmp/abv.err || { echo "[$1] FAILED"; tail -3 /tmp/abv.err; return; }
  python - "$1" <<'PY'
import json, sys
d = json.load(open("/tmp/abv.json"))
print("[%-10s] step %.3f ms  " % (sys.argv[1], d["ms_per_step"]) + "  ".join("%s %.3f" % (k[:22], v["ms"]) for k, v in d["roofline"]["kernels"].items()), flush=True)
PY
}
unset TIF_B200_LIB
run base
for n in "$@"; do
  export TIF_B200_LIB=$PWD/360opticalflow-tangentimages_b200/build/variants/$n.so
  run $n
done
