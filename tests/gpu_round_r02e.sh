# round-2 batch e: cluster / DSMEM exchange of the tcgen05 recurrent kernels: parity, phase timing, A/B against the L2 exchange
mkdir -p gpurun_out
O=gpurun_out/r02e
( timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
for ex in dsmem l2; do
  echo "== EXCHANGE=$ex" >> ${O}_timing.txt
  EESEN_B200_LSTM_EXCHANGE=$ex EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py >> ${O}_timing.txt 2>&1
done
for ex in dsmem l2 dsmem l2; do
  echo "== EXCHANGE=$ex" >> ${O}_ab.txt
  EESEN_B200_LSTM_EXCHANGE=$ex timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_timing.txt ${O}_ab.txt
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "bilstm_layer_vs_oracle or train_step" > ${O}_memcheck.log 2>&1; echo "memcheck exit $?" >> ${O}_memcheck.log )
tail -5 ${O}_memcheck.log
