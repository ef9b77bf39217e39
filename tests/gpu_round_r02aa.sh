mkdir -p gpurun_out
EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py > gpurun_out/r02aa_timing.txt 2>&1
cat gpurun_out/r02aa_timing.txt
