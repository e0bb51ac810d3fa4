#!/bin/bash
# Where does the wall time of one `blt` run go?  (stage log inside the library vs the whole process)
cd "$(dirname "$0")/.."
python - <<'PY'
import sys; sys.path.insert(0, ".")
from blt_b200 import synth
d = synth.text(1 << 30, 5); d.tofile("/dev/shm/x.bin")
l, r = synth.merges_from_sample(d, 60000); synth.write_merges_file("/dev/shm/m.txt", l, r)
PY
for i in 1 2 3; do
  s=$(date +%s%N)
  BLT_LOG=1 blt_b200/lib/blt -i /dev/shm/x.bin -o /dev/shm/y.bin --merges /dev/shm/m.txt --chunksize 16MB 2>&1 | grep -E "config parsed|device count|context created|buffers alloc|all shards"
  e=$(date +%s%N)
  echo "process wall $(( (e - s) / 1000000 )) ms"
done
rm -f /dev/shm/x.bin /dev/shm/y.bin /dev/shm/m.txt
