import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arm_spmv_b200 import host as H
torch.cuda.set_device(0)
A = H.uniform_coo(1 << 23, 1 << 23, 1 << 27, 43)
B = H.CSRMatrix(A); torch.cuda.synchronize(); del B
torch.cuda.cudart().cudaProfilerStart()
B = H.CSRMatrix(A); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
