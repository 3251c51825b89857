# Round-2 evidence run (one gpurun call, one GPU): tests, bench lines, launch list, --set full captures.
# Every ncu command follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2 > $O/r02_pytest.log
python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
python bench.py --workload large --no-extras --steps 5 --warmup 3 > $O/r02_bench_large_final.json 2> $O/r02_bench_large.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
# launch list of the default bench command (all sections)
python bench.py --steps 2 --warmup 3 > $O/p0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r02_launches_default.csv python bench.py --steps 2 --warmup 3 > $O/n0.log 2>&1
# C2: register-resident batched kernel
python bench.py --steps 2 --warmup 3 --no-extras > $O/p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:reg_simplex -s 2 -c 1 -f -o $O/prof_r02_reg python bench.py --steps 2 --warmup 3 --no-extras > $O/n1.log 2>&1
# C3: pass and look-ahead
python bench.py --workload large --no-extras --steps 2 --warmup 3 > $O/p2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream_update_pipe_tma -s 10 -c 1 -f -o $O/prof_r02_pass python bench.py --workload large --no-extras --steps 2 --warmup 3 > $O/n2.log 2>&1
# C4: condensed-tableau node kernel (control-warp design), a deep-node launch (one CTA per SM)
python tools/profile_targets.py bnb > $O/p3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cta_condensed -s 150 -c 2 -f -o $O/prof_r02_cond python tools/profile_targets.py bnb > $O/n3.log 2>&1
# C5: device-resident knapsack search (a persistent kernel: a small node pool keeps ncu's save / restore small)
LPX_KNAP_POOL_MB=1024 python tools/profile_targets.py knap > $O/p5.log 2>&1 && \
LPX_KNAP_POOL_MB=1024 ncu --set full --clock-control none --import-source on -k regex:knap_search -c 1 -f -o $O/prof_r02_knap python tools/profile_targets.py knap > $O/n5.log 2>&1
# Mode B: warm-started nodes
python tools/profile_targets.py pooled > $O/p6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cta_simplex_kernel -s 40 -c 1 -f -o $O/prof_r02_pooled python tools/profile_targets.py pooled > $O/n6.log 2>&1
cat $O/r02_pytest.log
for f in n0 n1 n2 n3 n5 n6; do tail -n 2 $O/$f.log; done
ls -la $O/*.ncu-rep
