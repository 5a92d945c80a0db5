#!/bin/bash
# Refreshes the evidence under profiles/ in one GPU box visit (run through gpurun; outputs land in gpurun_out/<tag>_*).
#   bench    : bench lines of cfg2 (default), cfg1, cfg4, cfg5
#   launches : the ncu launch list of the default bench command
#   ncu      : one ncu --set full capture of the three kernels of a step (FILTERED: an unfiltered capture of
#              tools/prof_step.py also replays every bank-registration and torch kernel ~40 times and once cost a
#              20-minute box visit and a report too large to be copied back)
tag=${1:-r4}; what=${2:-bench}
out=gpurun_out
case $what in
bench)
  timeout 240 python bench.py --steps 50 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "cfg2 exit $?"
  for c in cfg1 cfg4 cfg5; do timeout 200 python bench.py --config $c --steps 30 --warmup 5 > $out/${tag}_$c.json 2> $out/${tag}_$c.err; echo "$c exit $?"; done ;;
launches)
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_launches.log 2>&1; echo "launch list exit $?" ;;
ncu)
  timeout 120 python tools/prof_step.py --steps 3 > $out/${tag}_plain.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:conv_kernel|feat_frames_kernel|feat_epilogue' \
      --launch-skip 3 --launch-count 3 -o $out/${tag}_prof -f python tools/prof_step.py --steps 3 > $out/${tag}_ncu.log 2>&1; echo "ncu full exit $?" ;;
esac
