for v in t512b2 t256b4 t128b8 t256b3 t256b5 t256b6; do
  for n in 151552 227328; do
    echo "== $v n=$n $(AFF_LIB_PATH=/root/repo/build_variants/$v.so python tools/profile_run.py $n 2 'get fresh_tomatoes' sweep | grep best)"
  done
done
