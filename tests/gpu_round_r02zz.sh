# round-2 last check of the committed state
mkdir -p gpurun_out
O=gpurun_out/r02zz
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-alt > ${O}_bench.json 2> ${O}_bench.err
cut -c1-260 ${O}_bench.json
