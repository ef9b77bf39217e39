# round-2 batch h: scale double-buffer fix under the sanitizer, full GPU suite, bench + ncu captures of the cluster kernels
mkdir -p gpurun_out
O=gpurun_out/r02h
( timeout 120 eesen_b200/bin/umma_probe 2>&1 | grep "bwd step" ) > ${O}_umma_bwd.txt
for shape in "3 9 40 384" "3 9 40 256" "5 12 40 128" "3 9 40 320"; do
  echo "== memcheck $shape" >> ${O}_san.txt
  timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py $shape 2>&1 | grep -v "^=========\s*$" | grep "max\|elements\|ERROR SUMMARY\|Invalid" >> ${O}_san.txt
done
echo "== memcheck pytest layer tests (dsmem)" >> ${O}_san.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -k "bilstm_layer_vs_oracle and fp32x3" 2>&1 | grep "passed\|failed\|ERROR SUMMARY\|FAILED" >> ${O}_san.txt
( timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 14 -c 4 -o ${O}_lstm_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_ncu_full.log 2>&1
cat ${O}_umma_bwd.txt ${O}_san.txt; tail -3 ${O}_pytest.log; cat ${O}_bench.json | cut -c1-400
