# round-2 experiment batch: exchange / MMA micro-benchmarks, phase timing, stream-overlap A/B
mkdir -p gpurun_out
O=gpurun_out/r02c
( timeout 120 eesen_b200/bin/cluster_exchange2 ) > ${O}_cluster.txt 2>&1
( timeout 120 eesen_b200/bin/umma_probe ) > ${O}_umma.txt 2>&1
for sg in 1 0; do
  echo "== STREAM_GEMM=$sg" >> ${O}_timing.txt
  EESEN_B200_STREAM_GEMM=$sg EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py >> ${O}_timing.txt 2>&1
done
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== STREAM_GEMM=$1 OVERLAP=$2" >> ${O}_ab.txt
  EESEN_B200_STREAM_GEMM=$1 EESEN_B200_OVERLAP=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_cluster.txt ${O}_timing.txt ${O}_ab.txt; grep "timing TS" ${O}_umma.txt
