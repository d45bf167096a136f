#!/bin/bash
# usage: gr.sh <timeout_s> <gpus> '<command>'   -- retries while the pod answers busy / transient
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gr_last.out 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gr_last.out 2>&1; fi
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gr_last.out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -25 /tmp/gr_last.out
exit $rc
