#!/bin/bash
# Runs a command with GPU core dumps enabled and, if a kernel faults, prints which kernel / PC / exception (cuda-gdb reads the dump).
#   tools/coredump_run.sh <log> <command ...>
log=$1; shift
rm -f /tmp/b200vad_core_*
export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_ENABLE_LIGHTWEIGHT_COREDUMP=1 CUDA_COREDUMP_FILE=/tmp/b200vad_core_%p
"$@" > "$log" 2>&1
echo "rc=$?" >> "$log"
for c in /tmp/b200vad_core_*; do
  [ -f "$c" ] || continue
  echo "== GPU core dump $c ($(stat -c %s "$c") bytes)" >> "$log"
  /usr/local/cuda/bin/cuda-gdb-minimal -batch -ex "target cudacore $c" -ex "info cuda exception" -ex "info cuda kernels" -ex "bt" \
      -ex "info cuda lanes" -ex "x/6i \$pc-48" -ex "x/6i \$pc" >> "$log" 2>&1
done
