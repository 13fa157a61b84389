#!/bin/bash
# DRAM bytes and durations of every kernel of the settled 100k-box tick (plain launches instead of the graph, so that
# every kernel is listed); tools/wide_traffic_sum.py adds up the last tick.  Usage (under gpurun): bash tools/wide_traffic.sh <tag>
tag=${1:-r1}
GPX_WIDE_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_wide_traffic_all.csv python tools/wide_profile.py 100 10 100 30 > gpurun_out/${tag}_wide_traffic.log 2>&1
python tools/wide_traffic_sum.py gpurun_out/${tag}_wide_traffic_all.csv 43 > gpurun_out/${tag}_wide_traffic.txt
rm -f gpurun_out/${tag}_wide_traffic_all.csv
cat gpurun_out/${tag}_wide_traffic.txt
