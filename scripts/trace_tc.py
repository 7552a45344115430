import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyprob_b200 import _lib
from pyprob_b200._lib import call, ptr, stream
dev = torch.device('cuda:0')
def pack(x):
    rows, K = x.shape
    nfl = _lib.call('ppb_packed_floats', rows, K)
    hi = torch.empty(nfl, device=dev); lo = torch.empty(nfl, device=dev)
    call('ppb_pack_tf32', ptr(x), rows, K, x.stride(0), ptr(hi), ptr(lo), stream())
    return hi, lo
tr = torch.zeros(16, dtype=torch.int64, device=dev)
for (M, N, K, relu) in [(128, 128, 32, 0), (128, 128, 32, 8)]:
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); c = torch.empty(M, N, device=dev)
    ah, al = pack(a); bh, bl = pack(b)
    for it in range(3):
        call('ppb_debug_trace', ptr(tr))
        call('ppb_gemm_packed', ptr(ah), ptr(al), ptr(bh), ptr(bl), ptr(c), M, N, K, N, None, relu, 0, stream())
        torch.cuda.synchronize()
        t = tr.cpu().numpy()
        names = ['start', 'alloc', 'sync1', 'first_data', 'last_commit', 'acc_ready', 'stores_done', 'sync2', 'dealloc']
        print(M, N, K, 'relu', relu, ' '.join('%s=%d' % (n, t[i] - t[0]) for i, n in enumerate(names)))
    call('ppb_debug_trace', None)
