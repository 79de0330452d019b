# usage: bash tools/gpu_run.sh [tests] [phase] [modes...]
mkdir -p gpurun_out
for a in "$@"; do
  case $a in
    tests) timeout 600 python -m pytest tests -m gpu -q -x -k "pipe or golden" 2>&1 | tail -5 > gpurun_out/pytest_pipe.log; cat gpurun_out/pytest_pipe.log;;
    alltests) timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_all.log; cat gpurun_out/pytest_all.log;;
    phase) timeout 300 python tools/phase_times_pipe.py 8 > gpurun_out/phase_times.txt 2>&1; head -20 gpurun_out/phase_times.txt;;
    newtests) timeout 600 python -m pytest tests -m gpu -q -x -k "fill_uniform or abi" 2>&1 | tail -12 > gpurun_out/pytest_new.log; cat gpurun_out/pytest_new.log;;
    bwdtests) timeout 600 python -m pytest tests -m gpu -q -x -k "backward" 2>&1 | tail -12 > gpurun_out/pytest_bwd.log; cat gpurun_out/pytest_bwd.log;;
    bwd) (timeout 200 python tools/time_backward.py 8; timeout 200 python tools/time_backward.py 32 cam) 2>&1 | grep -v Warn > gpurun_out/time_backward.txt; cat gpurun_out/time_backward.txt;;
    bwd_ab) for l in nerf_from_image_b200/csrc/libnfi_render_*.so; do (NFI_LIB_PATH=$PWD/$l timeout 200 python tools/time_backward.py 8; NFI_LIB_PATH=$PWD/$l timeout 200 python tools/time_backward.py 32 cam) 2>&1 | grep -v Warn; done > gpurun_out/time_backward_ab.txt; cat gpurun_out/time_backward_ab.txt;;
    ncu_list) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-backward > gpurun_out/ncu_list.log 2>&1; tail -12 gpurun_out/launches.csv | cut -c1-220;;
    ncu_bwd) timeout 500 ncu --set full --clock-control none -k regex:render_backward_pipe -s 1 -c 1 -f -o gpurun_out/bwd_full python tools/time_backward.py 8 > gpurun_out/ncu_bwd.log 2>&1; tail -3 gpurun_out/ncu_bwd.log | cut -c1-200;;
    ncu_full) timeout 500 ncu --set full --clock-control none --import-source on -k regex:render_forward_pipe -s 2 -c 1 -f -o gpurun_out/pipe_full python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-backward > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200;;
    full) timeout 280 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -2 gpurun_out/bench_full.err;;
    *) timeout 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-backward --mlp-mode $a > gpurun_out/bench_m$a.json 2> gpurun_out/bench_m$a.err
       echo "mode $a: $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_m$a.json')); print(d['value']/1e6, 'Mrays/s kernel_ms', d['roofline']['kernel_ms'])" 2>&1 | tail -1)";;
  esac
done
