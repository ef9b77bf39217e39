# round-2 batch f: sanitizer runs on the C=384 layer test (failed once under memcheck with the cluster exchange)
mkdir -p gpurun_out
O=gpurun_out/r02f
K='bilstm_layer_vs_oracle and fp32x3-3-9-40-384'
for ex in dsmem l2; do
  for tool in memcheck initcheck; do
    echo "== EXCHANGE=$ex tool=$tool" >> ${O}_san.txt
    EESEN_B200_LSTM_EXCHANGE=$ex timeout 300 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep -v "^\s*$" | grep "passed\|failed\|ERROR SUMMARY\|Uninit\|Invalid\|=========     at\|AssertionError" | head -12 >> ${O}_san.txt
  done
done
for i in 1 2 3; do
  echo "== plain run $i (dsmem)" >> ${O}_san.txt
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | tail -1 >> ${O}_san.txt
done
echo "== racecheck dsmem" >> ${O}_san.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep "passed\|failed\|ERROR SUMMARY\|hazard\|=========     at" | head -20 >> ${O}_san.txt
cat ${O}_san.txt
