#!/bin/bash
# Round-2 closing run on one B200: staged parity tests, micro-benchmarks, racecheck of the attention kernels, the default
# bench line and smoke().
mkdir -p gpurun_out
o=gpurun_out
scripts/gpu_stage.sh r02_final > $o/r02_final_stage_tail.txt 2>&1
timeout 120 python scripts/gpu_norm_bench.py > $o/r02_norm_bench.txt 2>&1
timeout 120 python scripts/gpu_attn_bench.py > $o/r02_attn_bench.txt 2>&1
timeout 200 python scripts/gpu_skinny_bench.py --batches > $o/r02_skinny_batches.txt 2>&1
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_sanitizer_cases.py attn > $o/r02_racecheck_attn.txt 2>&1
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_sanitizer_cases.py attn decode > $o/r02_memcheck_attn_decode.txt 2>&1
timeout 900 python bench.py > $o/r02_bench_n1_final.json 2> $o/r02_bench_n1_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $o/r02_smoke.txt 2>&1
grep -E "^-----|passed|failed" $o/r02_final_stage.log | tail -30
cat $o/r02_norm_bench.txt $o/r02_attn_bench.txt
tail -2 $o/r02_racecheck_attn.txt $o/r02_memcheck_attn_decode.txt $o/r02_smoke.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'tok/s', d['value'], 'roofline', d['roofline']['frac'], d['roofline'].get('step_frac_of_peak'))
print('e2e', d['e2e']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print('decode', d['decode']['ms_per_step'], d['decode']['roofline']['frac'], d['decode'].get('batch32'))
print('clocks', d['clocks'])
PY
