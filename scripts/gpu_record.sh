#!/bin/bash
# usage (on the GPU box, from the repo root): scripts/gpu_record.sh <tag>
#   GPU test suite, the bench line, then the two ncu passes of B200_PROFILING.md (each only after the same
#   command exited 0 without ncu).  Everything lands in gpurun_out/; scripts/make_profiles.py turns it into profiles/.
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --batch 8 --no-e2e --no-cpu-baseline --no-extra-configs"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
# full capture of every hot kernel (scripts/prof_kernels.py: pass 1 = 8 matching launches to skip, pass 2 = the capture)
for h in 1024 1920; do
  CMD="python scripts/prof_kernels.py 8 $h"
  $CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on \
      -k regex:'k_erp2ico|k_lift|k_gather|k_blend|k_warp' -s 8 -c 8 -f -o gpurun_out/prof_${tag}_h$h $CMD > gpurun_out/ncu2_h$h.log 2>&1; echo "full capture h=$h rc=$?"
done
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("pairs/s %.0f  step %.3f ms  e2e %.0f (gray %.0f)  " % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["variants"]["gray"]["value"]) + "  ".join("%s %.3f" % (k[:22], v["ms"]) for k, v in d["roofline"]["kernels"].items()))
PY
