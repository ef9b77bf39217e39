mkdir -p gpurun_out
timeout 600 python tests/driver_throughput.py 4096 > gpurun_out/r02y_driver_throughput.json 2> gpurun_out/r02y_driver.err
cat gpurun_out/r02y_driver_throughput.json; tail -3 gpurun_out/r02y_driver.err
