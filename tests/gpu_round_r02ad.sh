mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_decoder.py -m gpu -x -q > gpurun_out/r02ad_decoder.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ad_decoder.log )
tail -15 gpurun_out/r02ad_decoder.log
