#!/bin/bash
# The throughput matrix north_star asks for, one reproducible script (run on the GPU box):
# bench.py's measurement at N = 65,536 / 262,144 / 1,048,576 on NGPU ranks, with both
# position-exchange mechanisms (1 = peer-memory stores fused into the step kernel, 0 = NCCL
# all-gather) when NGPU > 1, each line carrying the in-run parity record.  tools/matrix.py
# runs the cells of one GPU count in one process per rank (bench.py's own run_ours per cell).
# One JSON line per cell: OUT/<PREFIX>_scale_g<NGPU>_n<N>_x<EXCHANGE>.json
# usage: tools/run_scaling.sh NGPU [OUTDIR] [PREFIX] [SIZES]     e.g.  tools/run_scaling.sh 8 gpurun_out r2
# tools/report.py turns the lines (copied to profiles/) into profiles/<PREFIX>_summary.md.
NG=${1:-1}; OUT=${2:-gpurun_out}; PFX=${3:-r2}; SIZES=${4:-65536,262144,1048576}; mkdir -p $OUT
if [ "$NG" = 1 ]; then
  timeout 1200 python tools/matrix.py --gpus 1 --out $OUT --prefix $PFX --sizes $SIZES 2>&1 >/dev/null | grep '^gpus='
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 300)) tools/matrix.py --gpus $NG --out $OUT --prefix $PFX --sizes $SIZES 2>&1 >/dev/null | grep '^gpus='
fi
