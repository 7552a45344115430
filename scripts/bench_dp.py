"""Micro-benchmark of the data-parallel optimiser step at the BASELINE configs[1] arena size:
fused peer-memory kernel (ppb_dp_adam_step) vs NCCL all-reduce + device Adam.  Launch with torchrun."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyprob_b200 import parallel  # noqa: E402
from pyprob_b200._lib import call, ptr  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1721000
    iters = 200
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0 / world], device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    peer = parallel.PeerAdam(n, dev)
    peer.params.normal_()
    peer.grad.normal_()
    state = torch.zeros(4, dtype=torch.int32, device=dev)
    fused = timed(lambda: peer.step(m, v, hyper, state, st))
    trace = peer.phase_trace_us()
    assert not peer.timed_out()

    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    state2 = torch.zeros(4, dtype=torch.int32, device=dev)

    def unfused():
        dist.all_reduce(g)
        call('ppb_adam_step_dev', ptr(p), ptr(g), ptr(m), ptr(v), n, ptr(hyper), ptr(state2), st)
    nccl = timed(unfused)
    adam_only = timed(lambda: call('ppb_adam_step_dev', ptr(p), ptr(g), ptr(m), ptr(v), n, ptr(hyper), ptr(state2), st))
    if rank == 0:
        print(json.dumps({'world': world, 'floats': n, 'fused_us': fused, 'nccl_allreduce_plus_adam_us': nccl,
                          'adam_only_us': adam_only, 'fused_phase_us_rank0': trace}))
    peer.close()
    dist.barrier()
    os._exit(0)


if __name__ == '__main__':
    main()
