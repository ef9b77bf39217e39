mkdir -p gpurun_out
O=gpurun_out/r02m
for cfg in "0 0 1" "0 1 1" "1 1 1" "1 1 2" "1 1 3" "1 0 2"; do
  set -- $cfg
  echo "== STREAM_DX=$1 EARLY_CONV=$2 DX_READY=$3" >> ${O}_ab.txt
  EESEN_B200_STREAM_DX=$1 EESEN_B200_EARLY_CONV=$2 EESEN_B200_DX_READY=$3 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_ab.txt
