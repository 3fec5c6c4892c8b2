#!/bin/bash
# Multi-GPU evidence of one round on an N-GPU box:   gpurun --gpus N --timeout 1800 -- 'bash tools/run_multi.sh N r2'
# weak scaling (one process per GPU under torchrun, the driver's launch line) for C2 / C3 (/ C5 at N = 8), one
# sharded call over N GPUs (--strong), the sharded-call parity test, and the PCIe rate of all GPUs at once.
N=${1:-2}
R=${2:-r2}
O=gpurun_out
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p $O
nvidia-smi -L | head -8
run() {  # run <name> <torchrun? 1/0> <args...>
  local name=$1 tr=$2; shift 2
  if [ "$tr" = 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  else
    python bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  fi
  python - "$O/$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "value %.3g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.3g" % e.get("value", 0), "e2e ms %.2f" % e.get("ms_per_step", 0),
          "clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
except Exception as exc:
    print(sys.argv[1], "FAILED", exc)
PY
  tail -2 $O/$name.err | cut -c1-300
}
python -m pytest tests/test_sharded_call.py -m gpu -x -q 2>&1 | tail -2
run ${R}_bench_C2_${N}gpu 1 --steps 20 --warmup 5
run ${R}_bench_C2_strong_${N}gpu 0 --strong --steps 20 --warmup 5 --no-cpu-baseline
run ${R}_bench_C3_${N}gpu 1 --workload C3 --steps 5 --warmup 3
if [ "$N" = 8 ]; then
  run ${R}_bench_C5_8x125M 1 --workload C5 --steps 5 --warmup 3
  run ${R}_bench_C3_strong_8gpu 0 --strong --workload C3 --steps 5 --warmup 3 --no-cpu-baseline
fi
python exp/bw_all.py $N > $O/${R}_bw_all_${N}gpu.txt 2>&1; cat $O/${R}_bw_all_${N}gpu.txt
