# A/B the shared libraries under variants/ against the in-tree build (developer tool)
cp dynetlsm_b200/libdlsm.so /tmp/base.so
for v in base "$@"; do
  if [ $v != base ]; then cp variants/libdlsm_$v.so dynetlsm_b200/libdlsm.so; else cp /tmp/base.so dynetlsm_b200/libdlsm.so; fi
  echo "== $v"
  python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python tools/benchsum.py
  DLSM_TIMELINE=1 python tools/timeline_probe.py cfg2 1332 nointercepts 2>&1 | grep "hdp emission" | tail -1
done
cp /tmp/base.so dynetlsm_b200/libdlsm.so
