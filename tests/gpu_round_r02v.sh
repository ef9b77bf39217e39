# round-2 batch v: backward epilogue without the MMA-issuing warp (warp 4 takes both utterance halves)
mkdir -p gpurun_out
O=gpurun_out/r02v
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
for tn in 0 64 0 64; do
  echo "== TUNE=$tn" >> ${O}_ab.txt
  EESEN_B200_TUNE=$tn timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-alt --no-gpu-reference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step'])" >> ${O}_ab.txt 2>&1
done
cat ${O}_ab.txt
for tn in 0 64; do
  echo "== TUNE=$tn" >> ${O}_timing.txt
  EESEN_B200_TUNE=$tn EESEN_B200_LIB=$PWD/eesen_b200/lib_timing/libeesen_b200.so timeout 300 python tests/lstm_timing.py 2>&1 | sed -n '/backward/,$p' >> ${O}_timing.txt
done
cat ${O}_timing.txt
for shape in "3 9 40 384" "3 9 40 256" "3 9 40 320" "5 12 40 64"; do
  echo "== memcheck $shape" >> ${O}_san.txt
  timeout 300 compute-sanitizer --tool memcheck python tests/debug_cl384.py $shape 2>&1 | grep "max\|elements\|ERROR SUMMARY\|Invalid" >> ${O}_san.txt
done
cat ${O}_san.txt
