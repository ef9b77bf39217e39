mkdir -p gpurun_out
( timeout 120 eesen_b200/bin/umma_probe 2>&1 | grep "bwd step" ) > gpurun_out/r02ab_umma_noise.txt
cat gpurun_out/r02ab_umma_noise.txt
