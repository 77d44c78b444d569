#!/bin/bash
# Developer helper, run on the GPU box: bench.py at the three BASELINE sizes on
# NGPU ranks, plus config 5.  usage: tools/run_scaling.sh NGPU OUTDIR
NG=${1:-1}; OUT=${2:-gpurun_out}; mkdir -p $OUT
run() { if [ "$NG" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; fi; }
for B in 65536 262144 1048576; do
  K=5; [ $B = 65536 ] && K=50; [ $B = 262144 ] && K=20
  run bench.py --gpus $NG --steps $K --warmup 3 --bodies $B --no-cpu-baseline 2>/dev/null | grep '^{' > $OUT/scale_g${NG}_n${B}.json
  python - <<PY
import json
d=json.load(open("$OUT/scale_g${NG}_n${B}.json"))
print("gpus=%d n=%d pairs/s=%.4g ms/step=%.3f frac_nominal=%.3f e2e=%.4g fold=%.4g" % (d["n_gpus"], d["config"]["bodies"], d["value"], d["ms_per_step"], d["roofline"]["job_frac_of_nominal"], d["e2e"]["value"], d.get("equal_mass_path",{}).get("value",0)))
PY
done
run tools/config5.py --exchange 1 2>/dev/null | grep '^{' > $OUT/config5_g${NG}.json; cut -c1-600 $OUT/config5_g${NG}.json
