# round-2 batch ac: backward -- the other warps sleep while warp 0 issues the product
mkdir -p gpurun_out
O=gpurun_out/r02ac
for tn in 0 128 0 128; do
  echo "== TUNE=$tn" >> ${O}_ab.txt
  EESEN_B200_TUNE=$tn timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step']['lstm_bwd'], d['per_category_ms_per_step']['lstm_fwd'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_ab.txt
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
