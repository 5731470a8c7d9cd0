mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 19000 -c 330 --csv \
    --log-file gpurun_out/launches_r01d.csv python bench.py --steps 100 --warmup 3 --no-cpu --no-graph  \
    > gpurun_out/ncu_try.log 2>&1
echo rc=$?
wc -l gpurun_out/launches_r01d.csv; head -5 gpurun_out/launches_r01d.csv | cut -c1-300
grep -E "==(PROF|ERROR|WARNING)" gpurun_out/ncu_try.log | head
