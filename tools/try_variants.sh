#!/bin/bash
# Timing harness for kernel experiments: every build_variants/*.so is swapped in as the product library in turn.
cp affaction_b200/libaffpredict.so /tmp/keep.so
for v in build_variants/*.so; do
  cp $v affaction_b200/libaffpredict.so
  echo "== $v"
  python tools/profile_run.py ${N:-2368} 3 | grep "launch 2"
  [ -n "$QB" ] && python tools/quickbench.py 2>&1 | grep "n 32768"
done
cp /tmp/keep.so affaction_b200/libaffpredict.so
