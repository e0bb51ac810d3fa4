python - <<'PY'
import sys,os
sys.path.insert(0,'.')
from blt_b200 import synth
d=synth.text(4<<30, synth.SEED_CONFIG[5]); l,r=synth.merges_from_sample(d,60000); synth.write_merges_file('/dev/shm/m.txt',l,r); d.tofile('/dev/shm/in.bin')
PY
for i in 1 2; do BLT_LOG=1 blt_b200/lib/blt -i /dev/shm/in.bin -o /dev/shm/out.bin --merges /dev/shm/m.txt --chunksize 16MB --gpus 1; done
BLT_LOG=1 blt_b200/lib/blt -i /dev/shm/in.bin -o /dev/shm/out.bin --chunksize 16MB --gpus 1
rm -f /dev/shm/in.bin /dev/shm/out.bin /dev/shm/m.txt
