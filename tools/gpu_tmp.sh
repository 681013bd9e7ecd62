set -x
mkdir -p gpurun_out
rm -f gpurun_out/probe_wexp.log
cp tak_b200/lib/libtaknative.so /tmp/A.so
for rep in 1 2; do
for v in A E; do
  if [ $v = A ]; then cp /tmp/A.so tak_b200/lib/libtaknative.so; else cp build/dev/libtaknative_exp.so tak_b200/lib/libtaknative.so; fi
  echo "== variant $v (E: weight slabs fetched for the first ring revolution only; wrong results, timing only)" >> gpurun_out/probe_wexp.log
  timeout 200 python tools/probe_selfplay.py 6 4144 800 3 2>&1 | cut -c1-200 >> gpurun_out/probe_wexp.log
done
done
cp /tmp/A.so tak_b200/lib/libtaknative.so
cat gpurun_out/probe_wexp.log
