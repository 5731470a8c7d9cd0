#!/bin/bash
# round 2: the B = 1 drop-in step -- GPU parity suite, wall latency (RM3 and sphere shapes, full window) with the forces
# written straight into the pinned buffer (default) or through a copy node, per-kernel device times of the compact graph
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r02t_pytest.log
: > gpurun_out/r02t_b1.txt
for V in "1 1" "1 0" "0 1"; do
  set -- $V
  echo "HC_COMPACT_DIRECT=$1 HC_COMPACT_INLINE=$2" >> gpurun_out/r02t_b1.txt
  HC_COMPACT_DIRECT=$1 HC_COMPACT_INLINE=$2 python profiles/b1_probe.py rm3 2000 >> gpurun_out/r02t_b1.txt 2>&1
  HC_COMPACT_DIRECT=$1 HC_COMPACT_INLINE=$2 python profiles/b1_probe.py sphere 2000 >> gpurun_out/r02t_b1.txt 2>&1
done
cat gpurun_out/r02t_b1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18100 -c 40 --csv --log-file gpurun_out/r02t_b1_rm3_launches.csv python profiles/b1_probe.py rm3 30 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02t_b1_rm3_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:40], r[7], r[8], r[-1])
P
