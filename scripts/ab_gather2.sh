#!/bin/bash
# usage (2 GPUs): scripts/ab_gather2.sh "<defs A>" ...  -- N=2 bench with the fused gather for each set of -D knobs
for defs in "$@"; do
  TIF_NVCC_DEFS="$defs" python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
  TIF_NVCC_DEFS="$defs" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > /tmp/abg.json 2>/tmp/abg.err || tail -3 /tmp/abg.err
  python - "$defs" <<'PY'
import json, sys
d = json.load(open("/tmp/abg.json"))
g = d["gather"]
print("[%s] step %.3f ms  with gather %.3f ms (%.0f GB/s)  blend %.3f" % (sys.argv[1], d["ms_per_step"], g["ms_per_step"], g["rank0_ingress_gbs"], d["roofline"]["kernels"]["k_blend_nw"]["ms"]))
PY
done
python 360opticalflow-tangentimages_b200/_build.py --force > /dev/null
