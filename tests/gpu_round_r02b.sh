set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log )
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > gpurun_out/r02b_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 8 -c 4 -o gpurun_out/r02b_lstm_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > gpurun_out/r02b_ncu_full.log 2>&1
tail -3 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_bench.json
