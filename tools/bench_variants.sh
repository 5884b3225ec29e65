#!/bin/bash
# One GPU call, every opt-in variant: runs bench.py with the default path and with each result-preserving opt-in switched
# on (alone and all together) and collects the JSON lines in gpurun_out/variants.jsonl.
#   gpurun --timeout 1500 -- 'bash tools/bench_variants.sh 5 3'
STEPS=${1:-5}
WARMUP=${2:-3}
mkdir -p gpurun_out
OUT=gpurun_out/variants.jsonl
: > "$OUT"
run() {
  echo "== $*" >&2
  env "$@" python bench.py --steps "$STEPS" --warmup "$WARMUP" --no-cpu-baseline 2>>gpurun_out/variants.err | tail -1 >> "$OUT"
}
run AMB_VARIANT=default
run AMB_ORTHO_DOMINANCE=1
run AMB_DSM_BALANCED_GATHER=1
run AMB_DSM_STREAM_CHUNKS=4
run AMB_COMPACT_MIRRORS=1
run AMB_ORTHO_DOMINANCE=1 AMB_DSM_BALANCED_GATHER=1 AMB_DSM_STREAM_CHUNKS=4 AMB_COMPACT_MIRRORS=1
python - <<'PY'
import json
for ln in open("gpurun_out/variants.jsonl"):
    try:
        d = json.loads(ln)
    except Exception:
        print("unparsable:", ln[:120])
        continue
    c, r = d["config"], d["roofline"]["stage_ms"]
    rows, rest = c["grid"].split("x")
    cells = int(rows) * int(rest.split("@")[0])
    flags = [k for k in ("ortho_dominance_cull", "dsm_balanced_gather", "compact_mirrors") if c.get(k)]
    if c.get("dsm_stream_chunks", 1) > 1:
        flags.append("chunks=%d" % c["dsm_stream_chunks"])
    e2e_ms = 1e3 * cells / d["e2e"]["value"] if d.get("e2e") else float("nan")
    print("%-62s step %6.2f ms (bin %.2f gather %.2f ortho %.2f)  e2e %6.1f ms"
          % (",".join(flags) or "default", d["ms_per_step"], r["dsm_bin"], r["dsm_gather"], r["ortho"], e2e_ms))
PY
