mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pipe or golden" 2>&1 | tail -15 > gpurun_out/pytest_pipe.log; cat gpurun_out/pytest_pipe.log
timeout 300 python tools/phase_times_pipe.py 8 > gpurun_out/phase_times.txt 2>&1; head -12 gpurun_out/phase_times.txt
for m in 0x4 0x204 0x104; do
  timeout 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --mlp-mode $m > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err
  echo "mode $m: $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_m$m.json')); print(d['value']/1e6, 'Mrays/s kernel_ms', d['roofline']['kernel_ms'])" 2>&1 | tail -1)"
done
