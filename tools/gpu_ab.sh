#!/bin/bash
# A/B two builds of the library on the same box: tools/build/lib_old.so vs lib_new.so, alternating
mkdir -p gpurun_out
for rep in 1 2; do for v in old new; do
  DEBVADER_B200_LIB=$PWD/tools/build/lib_$v.so python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ab_${v}_$rep.json 2>/dev/null
done; done
python - <<'PY'
import json
for v in ("old","new"):
    for rep in (1,2):
        b=json.loads(open(f"gpurun_out/ab_{v}_{rep}.json").read().strip().splitlines()[-1])
        print(v, rep, round(b["value"]), round(b["ms_per_step"],3), " ".join("%s=%.2f"%(l["layer"][4:],l["ms"]) for l in b["layers"] if l["ms"]>0.4))
PY
