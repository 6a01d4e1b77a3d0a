import sys, types, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import torch, bench
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pb = bench.build_problem(dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "profile5"]), types.SimpleNamespace(kernel_path=1, retile=True), 0, 1, dev)
eng = pb["eng"]
for _ in range(2): eng.run()
torch.cuda.synchronize()
