#!/bin/bash
# round 2: one slice of the pass as the timed region launches it (ncu --set full), FP64 pipe microbenchmark with clocks / power
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rad_block -s 5000 -c 2 \
    -o gpurun_out/prof_radslice_r02a -f python bench.py --steps 60 --warmup 3 --no-cpu --no-b1 --no-parity --no-faithful-leg \
    > gpurun_out/ncu_radslice_r02a.log 2>&1
ls -la gpurun_out/prof_radslice_r02a.ncu-rep
cd profiles/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && cd ../..
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02_fp64_clocks.csv &
SMI=$!
sleep 1
./profiles/microbench/fp64_pipes > gpurun_out/r02_fp64_pipes.txt 2>&1
sleep 0.5
kill $SMI
cat gpurun_out/r02_fp64_pipes.txt
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_fp64_clocks.csv'))][1:]
sm=[float(r[1].split()[0]) for r in rows if len(r)>3]; pw=[float(r[3].split()[0]) for r in rows if len(r)>3]
print('clock samples %d: sm MHz min %.0f median %.0f max %.0f; power W max %.0f; reasons seen: %s' % (len(sm), min(sm), sorted(sm)[len(sm)//2], max(sm), max(pw), sorted(set(r[4].strip() for r in rows if len(r)>4))))
P
