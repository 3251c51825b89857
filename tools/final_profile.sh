set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -q -m gpu -x 2>&1 | tail -2 > gpurun_out/final_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --workload large --no-extras --steps 5 --warmup 3 > gpurun_out/bench_large_final.json 2> gpurun_out/bench_large_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/p1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_batched.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reg_simplex -s 2 -c 1 -f -o gpurun_out/prof_final_reg python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/n2.log 2>&1
python bench.py --workload large --no-extras --steps 2 --warmup 3 > gpurun_out/p2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_large.csv python bench.py --workload large --no-extras --steps 2 --warmup 3 > gpurun_out/n3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_update_pipe_tma -s 10 -c 2 -f -o gpurun_out/prof_final_pass python bench.py --workload large --no-extras --steps 2 --warmup 3 > gpurun_out/n4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_lookahead_pipe -s 10 -c 1 -f -o gpurun_out/prof_final_la python bench.py --workload large --no-extras --steps 2 --warmup 3 > gpurun_out/n5.log 2>&1
LPX_BNB_TRACE=1 python tools/gpu_probe.py bnbrep > gpurun_out/p3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cta_cluster_simplex -s 90 -c 1 -f -o gpurun_out/prof_final_cluster python tools/gpu_probe.py bnbrep > gpurun_out/n6.log 2>&1
cat gpurun_out/final_pytest.log
for f in n2 n4 n5 n6; do tail -n 2 gpurun_out/$f.log; done
