"""A/B timing of the opt-in kernels against the defaults, each arm in its own process (the switches are read from
the environment when the library / network handle is created):
  PPB_FUSED_CELL=1|2   synthetic 50-address training step (B=512), forward+backward+Adam, CUDA events
  PPB_MIXTURE_STAGED=1 mixture-of-Normals / mixture-of-TruncatedNormals log_prob at 2^24 particles, K=10
Prints one JSON object.  Timings only — correctness is the job of tests/test_fused_cell_gpu.py and
tests/test_scoring_staged_gpu.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_STEP = r'''
import sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, %r)
from pyprob_b200 import synthetic
from pyprob_b200._lib import call, ptr
from pyprob_b200.network import BatchStruct
from pyprob_b200.util import Optimizer
dev = torch.device('cuda:0'); rng = np.random.default_rng(0); B, T = 512, 50
net = synthetic.synthetic50_network(precision=0, T=T)
net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
net._create_optimizer(); net._sync_native()
enc = synthetic.synthetic50_batch(rng, B, T=T).encode(net)
grad = torch.zeros_like(net._arena.data)
img = torch.from_numpy(enc.pack().copy()).pin_memory(); dimg = img.to(dev)
bs = BatchStruct(); call('ppb_batch_from_image', img.data_ptr(), dimg.data_ptr(), img.numel(), C.byref(bs))
need = net._ensure_workspace(enc)
st = torch.cuda.current_stream().cuda_stream
loss = torch.empty((), device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
step_no = [0]
def step():
    grad.zero_()
    call('ppb_ic_loss_forward', net._handle, ptr(net._arena.data), C.byref(bs), ptr(net._workspace), need, 0, ptr(loss),
         ptr(status), None, 1, st)
    call('ppb_ic_loss_backward', net._handle, ptr(net._arena.data), ptr(grad), C.byref(bs), ptr(net._workspace), need, 0,
         1.0, st)
    step_no[0] += 1
    call('ppb_adam_step', ptr(net._arena.data), ptr(grad), ptr(net._exp_avg), ptr(net._exp_avg_sq), grad.numel(), 1e-3,
         0.9, 0.999, 1e-8, 0.0, step_no[0], 1.0, st)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print(json.dumps({'ms_per_step': e0.elapsed_time(e1) / 20, 'loss': float(loss)}))
''' % ROOT

_MIX = r'''
import sys, json, torch
sys.path.insert(0, %r)
from pyprob_b200 import ops
n, K = 1 << 24, 10
g = torch.Generator(device='cuda').manual_seed(0)
means = torch.randn(n, K, device='cuda', generator=g); sd = torch.rand(n, K, device='cuda', generator=g) + 0.2
probs = torch.rand(n, K, device='cuda', generator=g); v = torch.randn(n, device='cuda', generator=g)
lo, hi = v - 1.0, v + 1.0
lp = torch.empty(n, device='cuda')
out = {}
for name, fn in (('mixture_normal', lambda: ops.mixture_normal_log_prob(v, means, sd, probs, lp_out=lp)),
                 ('mixture_truncated_normal', lambda: ops.mixture_truncated_normal_log_prob(v, means, sd, probs, lo, hi, lp_out=lp))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    bytes_per = (3 * K + (2 if name == 'mixture_normal' else 4)) * 4
    out[name] = {'ms': ms, 'gbs': n * bytes_per / (ms * 1e-3) / 1e9, 'checksum': float(lp.double().sum())}
print(json.dumps(out))
''' % ROOT


def run(script, env_extra):
    env = dict(os.environ)
    for k in ('PPB_FUSED_CELL', 'PPB_MIXTURE_STAGED'):
        env.pop(k, None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, '-c', script], env=env, capture_output=True, text=True, timeout=180)
    if r.returncode != 0:
        return {'error': r.stderr.strip().splitlines()[-1] if r.stderr.strip() else 'failed'}
    return json.loads(r.stdout.strip().splitlines()[-1])


if __name__ == '__main__':
    res = {}
    if len(sys.argv) < 2 or sys.argv[1] != 'mix':
        res['synthetic50_step'] = {'default': run(_STEP, {}), 'fused_cell': run(_STEP, {'PPB_FUSED_CELL': '1'}),
                                   'persistent': run(_STEP, {'PPB_FUSED_CELL': '2'})}
    res['mixture_scoring'] = {'default': run(_MIX, {}), 'staged': run(_MIX, {'PPB_MIXTURE_STAGED': '1'})}
    print(json.dumps(res, indent=1))
