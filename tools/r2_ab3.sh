#!/bin/bash
# Interleaved A/B of environment settings on the training-step bench (the run-to-run spread of one setting is ~0.4 ms,
# larger than most effects): ROUNDS passes over all settings, 20 timed steps each, then min / median per setting.
#   ROUNDS=3 bash tools/r2_ab3.sh "VAR=a VAR2=b" "VAR=c" ...
ROUNDS=${ROUNDS:-3}
out=$(mktemp)
for r in $(seq 1 $ROUNDS); do
  i=0
  for setting in "$@"; do
    env $setting timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --skip-extras 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print($i, '%.3f' % d['ms_per_step'], d['clocks']['sm_mhz'], ','.join(d['clocks']['reasons']) or '-')" >> $out
    i=$((i+1))
  done
done
python - "$out" "$@" <<'PY'
import sys, statistics as st
rows = [l.split() for l in open(sys.argv[1])]
for i, name in enumerate(sys.argv[2:]):
    ms = [float(r[1]) for r in rows if int(r[0]) == i]
    mhz = [r[2] for r in rows if int(r[0]) == i]
    print(f"{name:45s} min {min(ms):.3f}  median {st.median(ms):.3f}  all {' '.join('%.2f' % m for m in ms)}  sm_mhz {' '.join(mhz)}")
PY
rm -f $out
