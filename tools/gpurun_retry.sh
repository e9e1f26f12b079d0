#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout-seconds> <command...>   — retries while gpurun answers "transient"/busy (exit 3)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc attempt=$i" >> $LOG
