cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name, env..., args
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --parity 0 $EXTRA > gpurun_out/r02f_$n.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02f_$n.json"))
print("$n value %.0f ms/step %.1f frac %.3f launches %d"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["gpu_launches"]))
PY
}
EXTRA="--pipeline 3" run p3 A=1
EXTRA="--pipeline 4" run p4 A=1
EXTRA="--pipeline 5" run p5 A=1
EXTRA="--pipeline 6" run p6 A=1
EXTRA="--pipeline 4" run p4_ph1_20_2 TC_PHASES=1,20,2
EXTRA="--pipeline 4" run p4_ph1_1 TC_PHASES=1,1
EXTRA="--pipeline 4" run p4_ph1_4 TC_PHASES=1,4
EXTRA="--pipeline 4" run p4_ph1_22_1 TC_PHASES=1,22,1
