#!/bin/bash
# Bare pinned-copy ceiling with N concurrent processes (one per GPU): bash tools/pcie_ceiling.sh N [--numa] > file
# Prints N JSON lines (one per GPU) followed by one summary line.
N=${1:-1}; shift
HERE=$(dirname "$0")
START=$(python3 -c "import time; print(time.time() + 4)")
for ((i = 0; i < N; i++)); do
  "$HERE/pcie_ceiling" --device $i --start-at $START "$@" > /tmp/pcie_ceiling_$i.json &
done
wait
cat /tmp/pcie_ceiling_*.json
python3 - "$N" <<'PY'
import glob, json, sys
rows = [json.loads(open(p).read()) for p in sorted(glob.glob("/tmp/pcie_ceiling_*.json"))][: int(sys.argv[1])]
worst = max(r["both_ms_mean"] for r in rows)
print(json.dumps({"summary": True, "ranks": len(rows), "step_ms_slowest_rank": worst,
                  "aggregate_gbs": sum(r["h2d_bytes"] + r["d2h_bytes"] for r in rows) / worst / 1e6,
                  "d2h_gbs_min_alone_in_time": min(r["d2h_gbs"] for r in rows), "numa": rows[0]["numa"]}))
PY
rm -f /tmp/pcie_ceiling_*.json
