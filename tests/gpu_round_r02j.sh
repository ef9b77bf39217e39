# round-2 batch j: shared-memory accesses as LDS/STS (address space kept), backward prefetch in front of the MMA issue
mkdir -p gpurun_out
O=gpurun_out/r02j
for tn in 0 32; do
  echo "== TUNE=$tn" >> ${O}_timing.txt
  EESEN_B200_TUNE=$tn EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py >> ${O}_timing.txt 2>&1
done
( timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
for tn in 0 32 0 32; do
  echo "== TUNE=$tn" >> ${O}_ab.txt
  EESEN_B200_TUNE=$tn timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_timing.txt ${O}_ab.txt; tail -3 ${O}_pytest.log
