set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launch.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:b2_fused -s 12 -c 3 -f -o gpurun_out/prof_r1_c2 $B > gpurun_out/ncu_c2.log 2>&1
C="python bench.py --config c3 --steps 1 --warmup 3"
$C > gpurun_out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:b2_fused -s 4 -c 1 -f -o gpurun_out/prof_r1_c3_argmax $C > gpurun_out/ncu_c3.log 2>&1
C="python bench.py --config c4 --steps 1 --warmup 3"
$C > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"b2_gather|b2_fused" -s 4 -c 3 -f -o gpurun_out/prof_r1_c4 $C > gpurun_out/ncu_c4.log 2>&1
C="python bench.py --config cum --steps 1 --warmup 3"
$C > gpurun_out/plain_cum.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:b2_fused -s 12 -c 3 -f -o gpurun_out/prof_r1_cum $C > gpurun_out/ncu_cum.log 2>&1
C="python bench.py --config c5 --steps 1"
$C > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o gpurun_out/prof_r1_c5_gemm $C > gpurun_out/ncu_c5.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for c in c1 c3 c4 c5 cum; do python bench.py --config $c --steps 10 > gpurun_out/bench_$c.jsonl 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log; cut -c1-200 gpurun_out/bench_c2.json; ls -la gpurun_out | tail -30
