#!/bin/bash
# The round's closing measurement set on one B200 (run through gpurun from the repo root): the full -m gpu suite, smoke(), the bench
# lines of the four BASELINE shapes + the reference arm, and the ncu launch list of the default bench command.
# Everything lands in gpurun_out/ (scratch); the summaries judged are copied into profiles/ by hand.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -x -q -m gpu) > gpurun_out/final_gpu_tests.log 2>&1; tail -3 gpurun_out/final_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
for W in c3 c4 c5; do timeout 300 python bench.py --steps 10 --warmup 3 --workload $W --no-cpu-baseline --e2e-modes bgzf+wire > gpurun_out/final_bench_$W.json 2> gpurun_out/final_bench_$W.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-modes bgzf+wire > gpurun_out/final_ncu.log 2>&1
for f in c2 c3 c4 c5 reference; do python - <<P
import json
try:
    d=json.loads(open("gpurun_out/final_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "clocks", d.get("clocks"))
except Exception as e:
    print("$f", "FAILED", e)
P
done
