#!/bin/bash
# The throughput matrix north_star asks for, one reproducible script (run on the GPU box):
# bench.py at N = 65,536 / 262,144 / 1,048,576 on NGPU ranks, with both position-exchange
# mechanisms (0 = NCCL all-gather, 1 = peer-memory stores fused into the step kernel) when
# NGPU > 1, each line carrying the in-run parity record.  One JSON line per cell:
#   OUT/<PREFIX>_scale_g<NGPU>_n<N>_x<EXCHANGE>.json
# usage: tools/run_scaling.sh NGPU [OUTDIR] [PREFIX]      e.g.  tools/run_scaling.sh 8 gpurun_out r2
# tools/report.py turns the lines (copied to profiles/) into profiles/<PREFIX>_summary.md.
NG=${1:-1}; OUT=${2:-gpurun_out}; PFX=${3:-r2}; mkdir -p $OUT
run() {
  if [ "$NG" = 1 ]; then timeout 600 python "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
         --master-port $((29600 + RANDOM % 300)) "$@"; fi
}
EXCH="1 0"; [ "$NG" = 1 ] && EXCH="1"
for B in 65536 262144 1048576; do
  K=5; [ $B = 65536 ] && K=200; [ $B = 262144 ] && K=40
  [ $B = 1048576 ] && [ "$NG" -ge 4 ] && K=10
  for X in $EXCH; do
    F=$OUT/${PFX}_scale_g${NG}_n${B}_x${X}.json
    run bench.py --gpus $NG --steps $K --warmup 3 --bodies $B --exchange $X --no-cpu-baseline 2>$OUT/${PFX}_scale_g${NG}_n${B}_x${X}.err | grep '^{' > $F
    python - <<PY
import json
try:
    d=json.load(open("$F"))
    p=d.get("parity") or {}
    print("gpus=%d n=%d x=%s pairs/s=%.4g ms/step=%.4f frac_nominal=%.3f e2e=%.4g (%.0f%% of resident) parity max=%.2g above=%s identical=%s" % (
        d["n_gpus"], d["config"]["bodies"], d["variant"]["exchange"], d["value"], d["ms_per_step"],
        d["roofline"]["job_frac_of_nominal"], d["e2e"]["value"], 100*d["e2e"]["value"]/d["value"],
        p.get("max_rel_vs_ref", float("nan")), p.get("n_above_1e-12"), p.get("ranks_bit_identical")))
except Exception as e:
    print("FAILED $F", e)
PY
  done
done
