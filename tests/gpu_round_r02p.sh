# round-2 batch p: end-of-round evidence -- full GPU suite, the driver's bench command, launch list, ncu --set full captures, C4 line
mkdir -p gpurun_out
O=gpurun_out/r02p
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 14 -c 4 -o ${O}_lstm_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc16 -s 150 -c 24 -o ${O}_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_ncu_gemm.log 2>&1
timeout 600 python bench.py --workload c4 --gemm-precision bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_bench_c4_bf16.json 2> ${O}_bench_c4.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > ${O}_bench_c4_fp32x3.json 2>> ${O}_bench_c4.err
cut -c1-300 ${O}_bench.json; cut -c1-200 ${O}_bench_c4_bf16.json; cut -c1-200 ${O}_bench_c4_fp32x3.json
