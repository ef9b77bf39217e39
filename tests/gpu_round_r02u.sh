# round-2 batch u: CTC branch-free log-sum-exp, prefetch distance 8
mkdir -p gpurun_out
O=gpurun_out/r02u
( timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > ${O}_bench.json 2> ${O}_bench.err
python -c "
import json
d=json.loads(open('${O}_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['per_category_ms_per_step']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])"
