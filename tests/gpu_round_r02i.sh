mkdir -p gpurun_out
O=gpurun_out/r02i
EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py > ${O}_timing.txt 2>&1
cat ${O}_timing.txt
