#!/bin/bash
# A/B of kernel builds on one box: every variants/*.so in turn takes the place of the product library,
# then the rates of the BASELINE configs (tools/config_rates.py) and the smoke check are run.
# The library that was in place is put back at the end.  usage (under gpurun): bash tools/kernel_ab.sh <tag>
TAG=${1:-ab}
LIB=robigo-luculenta_b200/librl_b200.so
cp $LIB /tmp/librl_b200.keep
mkdir -p gpurun_out
: > gpurun_out/${TAG}_ab.txt
for v in variants/*.so; do
  [ -f "$v" ] || continue
  cp $v $LIB; touch $LIB
  echo "== $v" | tee -a gpurun_out/${TAG}_ab.txt
  timeout 300 python tools/config_rates.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.txt
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.txt
done
cp /tmp/librl_b200.keep $LIB; touch $LIB
# the library in place under the environment settings of variants/env.txt (one per line)
if [ -f variants/env.txt ]; then
  while read -r setting; do
    [ -z "$setting" ] && continue
    echo "== in-place library, $setting" | tee -a gpurun_out/${TAG}_ab.txt
    env RL_RATES_ONLY=${RL_RATES_ONLY:-C2} $setting timeout 300 python tools/config_rates.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.txt
  done < variants/env.txt
fi
