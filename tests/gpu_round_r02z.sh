# round-2 final check: smoke(), full GPU suite, the driver's bench command (both arms)
mkdir -p gpurun_out
O=gpurun_out/r02z
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
cut -c1-260 ${O}_bench.json
