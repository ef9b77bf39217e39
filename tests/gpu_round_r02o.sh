# round-2 batch o: planes of m written by the forward kernel, FWD_READY=1 default
mkdir -p gpurun_out
O=gpurun_out/r02o
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
for i in 1 2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_ab.txt
rm -f ${O}_trace.txt
EESEN_B200_TRACE_FILE=${O}_trace.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > /dev/null 2>&1
python tests/trace_summary.py ${O}_trace.txt > ${O}_timeline.txt; cat ${O}_timeline.txt
