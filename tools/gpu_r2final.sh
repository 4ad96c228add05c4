#!/bin/bash
# Round 2, final 1-GPU pass: parity tests, smoke, every bench line, the reference arm, the ncu launch list of the default
# bench and one full capture of each decode kernel and of both encoders.   Usage: bash tools/gpu_r2final.sh <tag>
set -u
TAG=${1:-r02f}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_hca_decode.json 2> $OUT/${TAG}_bench_hca_decode.err
for w in adx_encode adx_decode hca_encode hca_decrypt hca_decrypt_decode; do
  timeout 600 python bench.py --workload $w --cpu-seconds 5 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err
done
for q in 2 3; do timeout 600 python bench.py --quality $q --no-cpu --no-companion > $OUT/${TAG}_bench_hca_decode_q$q.json 2>/dev/null; done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 300 python tools/bench_v3.py > $OUT/${TAG}_bench_v3.txt 2>&1; tail -2 $OUT/${TAG}_bench_v3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hca_|adx_|scatter_|gather_|pcm_' -c 400 --csv \
    --log-file $OUT/${TAG}_launches_hca_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_unpack_fast -s 4 -c 1 -o $OUT/${TAG}_prof_unpack -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-companion --e2e-steps 1 > $OUT/${TAG}_ncu_unpack.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_imdct_fast -s 4 -c 1 -o $OUT/${TAG}_prof_imdct -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-companion --e2e-steps 1 > $OUT/${TAG}_ncu_imdct.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hca_encode_kernel -s 4 -c 1 -o $OUT/${TAG}_prof_hca_encode -f \
    python bench.py --workload hca_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_encode.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:adx_encode_fast -s 4 -c 1 -o $OUT/${TAG}_prof_adx_encode -f \
    python bench.py --workload adx_encode --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/${TAG}_ncu_adx_encode.log 2>&1
python - <<PY
import json,glob
for p in sorted(glob.glob("$OUT/${TAG}_bench_*.json")):
    try:
        d=json.load(open(p)); r=d.get("roofline",{})
        print(p.split("${TAG}_bench_")[1], "value %.4g"%d["value"], "ms", round(d.get("ms_per_step",0),3), r.get("kernel"), r.get("frac"), "e2e", d.get("e2e",{}).get("ms_per_step"), "dev", d.get("e2e_device",{}).get("ms_per_step"), d.get("clocks",{}).get("reasons"))
    except Exception as e: print(p, "failed", e)
PY
