"""Config 3, device-resident: time of the whole pass with the coupling iterations finished (a) inside the last launch
behind the dense forecast (default: 6 compacted passes, stragglers first in launch C) and (b) entirely in compacted
passes (30 passes: launch C then holds no straggler), to separate the dense forecast from the stragglers' chain."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from roadsurf_b200 import lib
sys.argv = [sys.argv[0], "--workload", "c3"]
wl = bench.Workload(bench.parse_args(), 0, 1, torch.device("cuda", 0))
st = torch.cuda.current_stream()
res = {}
for passes in (6, 30, 0, 6):
    lib.set_option("coupling_compaction_passes", passes)
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); wl.step(st); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res["passes_%d" % passes] = round(min(ts[1:]), 2)
lib.set_option("coupling_compaction_passes", 6)
print(json.dumps(res))
