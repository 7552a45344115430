"""Saturating-size gate GEMM (4096 x 2048 x 512) through ppb_gemm_packed: PPB_PERSISTENT=0/1 set by the caller."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device('cuda:0')
print(os.environ.get('PPB_PERSISTENT', '0'), json.dumps({k: {a: round(b, 4) for a, b in v.items()} for k, v in bench.gate_gemm_saturating(dev, bench.measured_peaks()).items()}))
