mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ae_pytest.log )
tail -3 gpurun_out/r02ae_pytest.log
