mkdir -p gpurun_out
O=gpurun_out/r02g
echo "== plain 3 9 40 384" > ${O}_dbg.txt
timeout 200 python tests/debug_cl384.py 3 9 40 384 >> ${O}_dbg.txt 2>&1
echo "== memcheck 3 9 40 384" >> ${O}_dbg.txt
timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py 3 9 40 384 2>&1 | grep -v "^=========" >> ${O}_dbg.txt
echo "== memcheck 16 9 40 384" >> ${O}_dbg.txt
timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py 16 9 40 384 2>&1 | grep -v "^=========" >> ${O}_dbg.txt
echo "== memcheck 3 9 40 320" >> ${O}_dbg.txt
timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py 3 9 40 320 2>&1 | grep -v "^=========" >> ${O}_dbg.txt
echo "== memcheck 3 9 40 256" >> ${O}_dbg.txt
timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py 3 9 40 256 2>&1 | grep -v "^=========" >> ${O}_dbg.txt
for tn in 0 16; do
  echo "== timing TUNE=$tn" >> ${O}_dbg.txt
  EESEN_B200_TUNE=$tn EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py 2>&1 | sed -n '/backward/,$p' >> ${O}_dbg.txt
done
cat ${O}_dbg.txt
