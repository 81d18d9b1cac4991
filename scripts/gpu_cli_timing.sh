python - <<'PY'
import os,sys
sys.path.insert(0,'.')
from dashing2_b200 import synth
paths = synth.write_fasta_set('/dev/shm/fa', 256, 5000000, seed=2, n_families=32)
open('/dev/shm/files.txt','w').write("\n".join(paths)+"\n")
PY
for i in 1 2; do dashing2_b200/bin/dashing2-gpu sketch -v -k31 -p16 -F /dev/shm/files.txt -o /dev/shm/o.stk --binary-output --cmpout /dev/shm/o.f32 -S1024 2>&1 | grep dashing2-gpu; echo; done
dashing2_b200/bin/dashing2-gpu sketch -v -k31 -p16 -F /dev/shm/files.txt -o /dev/shm/o.stk --binary-output --cmpout /dev/shm/o.f32 -S4096 -w51 --full-setsketch 2>&1 | grep dashing2-gpu
