#!/bin/bash
# One GPU-box visit: parity tests, the bench (both arms), then the ncu launch list and full captures of the hot kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
python bench.py --impl reference --steps 100 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?" >> $out/${tag}_bench.err
python tools/phase_profile.py 4096 20 > $out/${tag}_phases.txt 2>&1
# launch list of the headline step alone (value + e2e loops), then of the whole bench's first 1500 launches
P="python bench.py --steps 12 --warmup 3 --no-cpu --headline-only"
$P > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv $P > $out/${tag}_ncu1.log 2>&1
Q="python bench.py --steps 12 --warmup 3 --ray-reps 3 --no-wide --no-cpu"
$Q > $out/${tag}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tick -s 20 -c 2 -o $out/${tag}_tick $Q > $out/${tag}_ncu2.log 2>&1
$Q > $out/${tag}_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_raycast -s 20 -c 1 -o $out/${tag}_rays $Q > $out/${tag}_ncu3.log 2>&1
# launch list of the wide tick (plain launches instead of the graph, so every kernel is listed; the last ticks are settled)
python tools/wide_profile.py 100 10 100 60 > $out/${tag}_wide.txt 2>&1 && \
GPX_WIDE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_wide_launches_all.csv python tools/wide_profile.py 100 10 100 30 > $out/${tag}_ncu4.log 2>&1
tail -n 90 $out/${tag}_wide_launches_all.csv > $out/${tag}_wide_launches.csv; rm -f $out/${tag}_wide_launches_all.csv
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench.json | cut -c1-600; tail -2 $out/${tag}_bench.err
