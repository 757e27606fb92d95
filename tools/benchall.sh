for c in "2 125000" "3 125000" "5 25000" "4 128" "4 16" "1 100000"; do set -- $c; timeout 300 python bench.py --config $1 --pairs-per-gpu $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_cfg$1_$2.json 2> gpurun_out/b_cfg$1_$2.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_cfg$1_$2.json")); print("cfg$1 n=$2", round(d["value"],1), round(d["e2e"]["value"],1), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["breakdown_ms_per_step"].items()})
except Exception as ex: print("cfg$1 failed", ex)
PY
done
