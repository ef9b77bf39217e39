# round-2 batch r (2 GPUs): N-GPU == 1-GPU equivalence tests and the driver's N = 2 bench command
mkdir -p gpurun_out
O=gpurun_out/r02r
nvidia-smi -L > ${O}_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log )
tail -3 ${O}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench2.json 2> ${O}_bench2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-alt > ${O}_bench1.json 2> ${O}_bench1.err
python -c "
import json
for f in ('${O}_bench1.json','${O}_bench2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d.get('replicas_identical'), d['per_category_ms_per_step'])"
tail -3 ${O}_bench2.err
