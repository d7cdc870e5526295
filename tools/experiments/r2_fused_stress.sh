# round 2: is the (50 us) fused exit reliable with three batches in flight?  N short benches in a row, each under a timeout.
N=${1:-8}; shift
ok=0; bad=0
for i in $(seq 1 $N); do
  env "$@" timeout -s KILL 60 python bench.py --verbose --steps 250 --warmup 3 --skip-cpu > gpurun_out/st_$i.json 2> gpurun_out/st_$i.err
  rc=$?
  if [ $rc -eq 0 ]; then ok=$((ok+1)); else bad=$((bad+1)); echo "run $i rc=$rc last marker: $(tail -1 gpurun_out/st_$i.err)"; fi
done
echo "ok=$ok bad=$bad  ($*)"
python - <<PY
import json,glob
v=[json.loads(open(f).read().strip().splitlines()[-1])["value"] for f in sorted(glob.glob("gpurun_out/st_*.json")) if open(f).read().strip()]
print("clips/s:", [round(x) for x in v])
PY
