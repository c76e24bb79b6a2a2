# A/B the shared libraries under variants/ against the in-tree build (developer tool)
# usage: bash tools/ab_variants.sh <variant> [<variant> ...]   (variants/libdlsm_<variant>.so)
cp dynetlsm_b200/libdlsm.so /tmp/base.so
for v in base "$@"; do
  if [ $v != base ]; then cp variants/libdlsm_$v.so dynetlsm_b200/libdlsm.so; else cp /tmp/base.so dynetlsm_b200/libdlsm.so; fi
  echo "== $v"
  for w in cfg2 cfg4; do python bench.py --workload $w --steps 60 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python tools/benchsum.py; done
done
cp /tmp/base.so dynetlsm_b200/libdlsm.so
