#!/bin/bash
# measurement record: 20-step bench line + ncu launch list of the last steady-state step (first half of gpu_profile.sh)
set -o pipefail
TAG=${1:-rec}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_bench_$TAG.log 2>&1
python - <<PY
rows = [l for l in open("gpurun_out/launches_all_$TAG.csv") if l.startswith('"')]
hdr, body = rows[0], rows[1:]
idx = [i for i, l in enumerate(body) if "warp_crop_kernel" in l]
last = body[idx[-1]:]
open("gpurun_out/launches_$TAG.csv", "w").write(hdr + "".join(last))
print("launch list: %d launches in total, last step has %d" % (len(body), len(last)))
PY
rm -f gpurun_out/launches_all_$TAG.csv
python profiles/summarize_launches.py gpurun_out/launches_$TAG.csv --by-grid > gpurun_out/launches_$TAG.md; head -30 gpurun_out/launches_$TAG.md
