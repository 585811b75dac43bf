"""The bench's config-3 workload (10^5 coupled points from hourly records), two passes, for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rs_run_kernel|partition_kernel|rs_solar" -s <first pass> ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
args = bench.parse_args.__wrapped__() if hasattr(bench.parse_args, "__wrapped__") else None
sys.argv = [sys.argv[0], "--workload", "c3"]
args = bench.parse_args()
wl = bench.Workload(args, 0, 1, torch.device("cuda", 0))
st = torch.cuda.current_stream()
for _ in range(2):
    wl.step(st)
torch.cuda.synchronize()
print("ok", wl.P, wl.sim_len, wl.window_end, wl.lib.last_launch())
