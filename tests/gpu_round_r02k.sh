mkdir -p gpurun_out
rm -f gpurun_out/r02k_trace.txt
EESEN_B200_TRACE_FILE=gpurun_out/r02k_trace.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference > gpurun_out/r02k_bench.json 2>/dev/null
python tests/trace_summary.py gpurun_out/r02k_trace.txt > gpurun_out/r02k_timeline.txt
cat gpurun_out/r02k_timeline.txt
