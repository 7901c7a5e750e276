#!/bin/bash
# usage: tune_sweep.sh <workload> [extra bench args] -- "<tune set 1>" "<tune set 2>" ...   (a tune set = "k=v k=v")
w=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for t in "$@"; do
  args=()
  for kv in $t; do args+=(--tune "$kv"); done
  python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline "${extra[@]}" "${args[@]}" > gpurun_out/b.json 2>/dev/null
  python - "$w" "$t" <<'PY'
import json, sys
d = json.load(open("gpurun_out/b.json"))
r = d["roofline"]
print(sys.argv[1], "[%s]" % sys.argv[2], "seg/s", round(d["value"]), "ms", round(d["ms_per_step"], 1), "align", round(d["phase_ms_per_step"]["align"], 1),
      "frac", round(r["frac"], 3), "exec", round(r["executed_cell_fraction"], 4), "retries", r["band_retries_per_step"], "pairs", d["pairs_launched_per_step"])
PY
done
