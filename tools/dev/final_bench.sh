#!/bin/bash
# dev: the driver's command (--steps 20 --warmup 5) and the default (200 / 20) at N GPUs; JSON lines land in gpurun_out/
N=${1:-1}
run() {  # steps warmup tag
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps $1 --warmup $2 > gpurun_out/final_n${N}_$3.json 2> gpurun_out/final_n${N}_$3.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps $1 --warmup $2 > gpurun_out/final_n${N}_$3.json 2> gpurun_out/final_n${N}_$3.err
  fi
  echo "N=$N $3 rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/final_n${N}_$3.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"]*1e3,2), "us", round(d["value"]/1e3,1), "G/s frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None,
      "clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "stanh", d["stanh_step"]["in_flight"] if d.get("stanh_step") else None,
      "match", (d.get("exchange_check") or {}).get("match"), {k: (round(v["ms_per_step"]*1e3,2), (v.get("exchange_check") or {}).get("match")) for k, v in d["per_config"].items()})
PY
}
run 20 5 steps20
run 200 20 steps200
