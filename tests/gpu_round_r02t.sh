mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_kernel -s 3 -c 1 -o gpurun_out/r02t_ctc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > gpurun_out/r02t_ncu.log 2>&1
tail -2 gpurun_out/r02t_ncu.log
