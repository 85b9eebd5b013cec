python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair_cluster.py tests/test_gpu_reference_goldens.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
python - <<'P'
import torch, time, numpy as np
from admp_b200 import workloads
from admp_b200.neighbor import neighbor_list
import os
for name, w, rc in (('C2', workloads.water_box((1,1,1), polarizable=True), 4.0), ('dense32', workloads.dense_water(32), 8.0), ('C5', workloads.water_box((4,8,8), polarizable=True), 4.0)):
    nb = neighbor_list(w.box, rc).allocate(w.positions)
    pos = torch.as_tensor(w.positions, device='cuda')
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = nb.update(pos)
    a.record()
    for _ in range(20): nb = nb.update(pos)
    b.record(); b.synchronize()
    print(name, 'atoms', w.n_atoms, 'pairs', int(nb.n_pairs), 'nblist ms', a.elapsed_time(b)/20, os.environ.get('ADMP_NBLIST'))
P
ADMP_NBLIST=thread python - <<'P'
import torch, time, numpy as np
from admp_b200 import workloads
from admp_b200.neighbor import neighbor_list
import os
for name, w, rc in (('C2', workloads.water_box((1,1,1), polarizable=True), 4.0), ('dense32', workloads.dense_water(32), 8.0), ('C5', workloads.water_box((4,8,8), polarizable=True), 4.0)):
    nb = neighbor_list(w.box, rc).allocate(w.positions)
    pos = torch.as_tensor(w.positions, device='cuda')
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = nb.update(pos)
    a.record()
    for _ in range(20): nb = nb.update(pos)
    b.record(); b.synchronize()
    print(name, 'atoms', w.n_atoms, 'pairs', int(nb.n_pairs), 'nblist ms', a.elapsed_time(b)/20, os.environ.get('ADMP_NBLIST'))
P
