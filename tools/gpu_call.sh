# Round-2 GPU jobs (run on the box through gpurun); everything lands in gpurun_out/.
# usage: bash tools/gpu_call.sh <job> [...]
mkdir -p gpurun_out
LIBDIR=$PWD/nerf_from_image_b200/csrc
for a in "$@"; do
  case $a in
    tests) timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log;;
    bench) timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err;;
    ab) for l in $LIBDIR/libnfi_render.so $LIBDIR/libnfi_render_*.so; do NFI_LIB_PATH=$l timeout 120 python tools/ab_forward.py 32 10 2>&1 | grep -v Warn | tail -1; done > gpurun_out/r2_ab_forward.txt; cat gpurun_out/r2_ab_forward.txt;;
    ab_bwd) for l in $LIBDIR/libnfi_render.so $LIBDIR/libnfi_render_*.so; do NFI_LIB_PATH=$l timeout 200 python tools/time_backward.py 32 cam 2>&1 | grep -v Warn; NFI_LIB_PATH=$l timeout 200 python tools/time_backward.py 32 2>&1 | grep -v Warn | tail -1; done > gpurun_out/r2_ab_backward.txt; cat gpurun_out/r2_ab_backward.txt;;
    ncu_fwd) timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_forward_pipe -s 3 -c 1 -f -o gpurun_out/r2_fwd python tools/ab_forward.py 32 2 > gpurun_out/r2_ncu_fwd.log 2>&1
             ncu -i gpurun_out/r2_fwd.ncu-rep --page raw --csv > gpurun_out/r2_fwd_raw.csv 2>/dev/null
             ncu -i gpurun_out/r2_fwd.ncu-rep --page source --csv > gpurun_out/r2_fwd_src.csv 2>/dev/null
             rm -f gpurun_out/r2_fwd.ncu-rep; tail -2 gpurun_out/r2_ncu_fwd.log | cut -c1-200;;
    ncu_bwd) timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_backward_pipe -s 2 -c 1 -f -o gpurun_out/r2_bwd python tools/time_backward.py 32 cam > gpurun_out/r2_ncu_bwd.log 2>&1
             ncu -i gpurun_out/r2_bwd.ncu-rep --page raw --csv > gpurun_out/r2_bwd_raw.csv 2>/dev/null
             ncu -i gpurun_out/r2_bwd.ncu-rep --page source --csv > gpurun_out/r2_bwd_src.csv 2>/dev/null
             rm -f gpurun_out/r2_bwd.ncu-rep; tail -2 gpurun_out/r2_ncu_bwd.log | cut -c1-200;;
    ncu_list) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_list.log 2>&1; tail -14 gpurun_out/r2_launches.csv | cut -c1-200;;
    synth) timeout 900 python -m pytest tests/test_synthesis_gpu.py -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2_pytest_synth.log; cat gpurun_out/r2_pytest_synth.log;;
    diag2) (timeout 300 python tools/grad_diag2.py p3d_plain; timeout 300 python tools/grad_diag2.py cub_ortho; timeout 300 python tools/grad_diag2.py p3d_bbox) > gpurun_out/r2_grad_diag2.txt 2>&1; cat gpurun_out/r2_grad_diag2.txt;;
    tsynth) (timeout 400 python tools/time_synthesis.py 32 5; timeout 300 python tools/time_synthesis.py 4 5) > gpurun_out/r2_time_synthesis.txt 2>&1; cat gpurun_out/r2_time_synthesis.txt;;
    ncu_synth) timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 3 -f -o gpurun_out/r2_synth python tools/time_synthesis.py 8 1 > gpurun_out/r2_ncu_synth.log 2>&1
             ncu -i gpurun_out/r2_synth.ncu-rep --page raw --csv > gpurun_out/r2_synth_raw.csv 2>/dev/null
             rm -f gpurun_out/r2_synth.ncu-rep; tail -2 gpurun_out/r2_ncu_synth.log | cut -c1-200;;
    list_synth) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_synth_launches.csv python tools/time_synthesis.py 32 1 > gpurun_out/r2_list_synth.log 2>&1; python tools/launch_table.py gpurun_out/r2_synth_launches.csv | tail -40;;
    heads) timeout 600 python -m pytest tests/test_heads_gpu.py -m gpu -q --maxfail=12 2>&1 | tail -40 > gpurun_out/r2_pytest_heads.log; cat gpurun_out/r2_pytest_heads.log;;
    mgpu2|mgpu4|mgpu8) n=${a#mgpu}; TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
             (timeout 400 $TR bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/r2_bench_weak_n$n.json) 2> gpurun_out/r2_bench_weak_n$n.err
             (timeout 300 $TR bench.py --gpus $n --scaling strong --quick --no-e2e | tail -1 > gpurun_out/r2_bench_strong_n$n.json) 2> gpurun_out/r2_bench_strong_n$n.err
             (timeout 300 $TR bench.py --gpus $n --exchange nccl --quick --no-e2e | tail -1 > gpurun_out/r2_bench_nccl_n$n.json) 2> gpurun_out/r2_bench_nccl_n$n.err
             (timeout 300 $TR bench.py --gpus $n --exchange nccl --scaling strong --quick --no-e2e | tail -1 > gpurun_out/r2_bench_ncclstrong_n$n.json) 2> gpurun_out/r2_bench_ncclstrong_n$n.err
             (timeout 300 $TR bench.py --gpus $n --config 3 | tail -1 > gpurun_out/r2_bench_config3_n$n.json) 2> gpurun_out/r2_bench_config3_n$n.err
             (timeout 300 $TR bench.py --gpus $n --config 4 | tail -1 > gpurun_out/r2_bench_config4_n$n.json) 2> gpurun_out/r2_bench_config4_n$n.err
             for f in weak strong nccl ncclstrong config3 config4; do echo "== $f n=$n"; cut -c1-700 gpurun_out/r2_bench_${f}_n$n.json; tail -3 gpurun_out/r2_bench_${f}_n$n.err | cut -c1-300; done;;
    cfg1) (timeout 300 python bench.py --config 3 | tail -1 > gpurun_out/r2_bench_config3_n1.json) 2> gpurun_out/r2_bench_config3_n1.err; (timeout 300 python bench.py --config 4 | tail -1 > gpurun_out/r2_bench_config4_n1.json) 2> gpurun_out/r2_bench_config4_n1.err; (timeout 600 python bench.py --config 5 | tail -1 > gpurun_out/r2_bench_config5_n1.json) 2> gpurun_out/r2_bench_config5_n1.err
          for f in config3 config4 config5; do echo "== $f"; cut -c1-1500 gpurun_out/r2_bench_${f}_n1.json; tail -3 gpurun_out/r2_bench_${f}_n1.err | cut -c1-300; done;;
    ncu_normals) timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_normals_pipe -s 2 -c 1 -f -o gpurun_out/r2_normals python tools/time_normals.py 32 > gpurun_out/r2_ncu_normals.log 2>&1
             ncu -i gpurun_out/r2_normals.ncu-rep --page raw --csv > gpurun_out/r2_normals_raw.csv 2>/dev/null
             rm -f gpurun_out/r2_normals.ncu-rep; tail -2 gpurun_out/r2_ncu_normals.log | cut -c1-200;;
    smoke) timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3;;
    ncu_wgrad) timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wgrad_pipe -s 6 -c 1 -f -o gpurun_out/r2_wgrad python tools/time_wgrad.py 32 > gpurun_out/r2_ncu_wgrad.log 2>&1
             ncu -i gpurun_out/r2_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_wgrad_raw.csv 2>/dev/null
             rm -f gpurun_out/r2_wgrad.ncu-rep; tail -2 gpurun_out/r2_ncu_wgrad.log | cut -c1-200;;
    list_cfg4) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_config4.csv python bench.py --config 4 --steps 2 --warmup 3 > gpurun_out/r2_list_cfg4.log 2>&1; python tools/launch_table.py gpurun_out/r2_launches_config4.csv | tail -12;;
    diag) timeout 600 python tools/grad_diag.py > gpurun_out/r2_grad_diag.txt 2>&1; cat gpurun_out/r2_grad_diag.txt;;
    *) echo "unknown job $a";;
  esac
done
