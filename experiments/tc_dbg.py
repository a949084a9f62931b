import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keops_b200.codegen import knn_tc
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
D, K = 64, 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
x, y = torch.randn(M, D, device=dev), torch.randn(N, D, device=dev)
out = torch.empty(M, K, device=dev)
eng = knn_tc.KnnTensorCore(D, K, "ArgKMin", True, 0)
os.environ["KT_TIME"] = "1"
for _ in range(2):
    eng.run(x, y, out, torch.cuda.current_stream())
