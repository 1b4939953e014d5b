for v in s0x37 s0x2f s0x27 s0x2b s0x15; do
  echo "== $v $(AFF_LIB_PATH=/root/repo/build_variants/$v.so python tools/profile_run.py 151552 2 'get fresh_tomatoes' sweep | grep best)"
done
