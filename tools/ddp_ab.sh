# A/B of the data-parallel knobs (bench.py under torchrun): usage: bash tools/ddp_ab.sh N name ENV=... [ENV=...] -- name2 ENV=... 
N=$1; shift
run() {
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/ab${N}_$name.json 2> gpurun_out/ab${N}_$name.err
  python - <<PY
import json
for l in open("gpurun_out/ab${N}_$name.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N $name", round(d["ms_per_step"],3), round(d["value"]))
PY
}
args=()
for a in "$@"; do
  if [ "$a" == "--" ]; then run "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run "${args[@]}"
