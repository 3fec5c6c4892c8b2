set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests/test_sharded_call.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -15 gpurun_out/r2d_pytest.log
python bench.py --strong --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_strong2.json 2> gpurun_out/r2d_strong2.err; tail -c 1500 gpurun_out/r2d_strong2.json; tail -5 gpurun_out/r2d_strong2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_weak2.json 2> gpurun_out/r2d_weak2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_weak2.json').read().strip().splitlines()[-1])
print('weak2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('e2e_pinned',{}).get('ms_per_step'))
PY
tail -3 gpurun_out/r2d_weak2.err
