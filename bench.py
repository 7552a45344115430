#!/usr/bin/env python
"""bench.py — one JSON line per run (driver contract, tier framing (4)).

Workload (N = 1): BASELINE.json configs[1] — GaussianUnknownMean inference compilation,
InferenceNetworkLSTM h=512, observe embeddings 32+32, minibatch 256 prior traces, one Adam step per batch.
A "step" = one pass of the hot path over one synthetic minibatch: encode image -> forward -> hand-written
backward -> (N>1: one NCCL all-reduce of the flat gradient arena) -> fused Adam.

  value : traces/s, whole job, batch image resident in HBM when the timed region starts (device events,
          L2 flushed between steps, max over ranks)
  e2e   : same metric through the C-ABI host-buffer call (ppb_ic_train_step_host): pinned host image ->
          H2D -> forward/backward/Adam -> D2H loss, every step
  roofline     : LSTM gate GEMM class (input projections + recurrent GEMMs, fwd+bwd), tensor-core bound
  cpu_baseline : the oracle port of the reference's _loss + backward + Adam on this box's host cores
  extra        : IS / IC posterior particles/s through the public Model API (the metric's other half)

`--impl reference` times the CPU oracle port (the reference cannot travel to the GPU box) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 256
LSTM_DIM = 512
WORKLOAD = 'GaussianUnknownMean IC train, LSTM h=512, obs-embed 32+32, batch 256/GPU (BASELINE configs[1])'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The region lasts tens of milliseconds, far
    below nvidia-smi's loop period, so NVML is polled directly every ~2 ms from a thread; `nvidia-smi -lms` is
    the fallback when the NVML binding is missing."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NVML_REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'),
                    (0x4, 'sw_power_cap'), (0x80, 'hw_power_brake_slowdown'))

    def __init__(self, gpu_index, uuid=None):
        self.rows, self.proc, self.gpu, self.uuid = [], None, gpu_index, uuid
        self.nvml, self.handle, self.thread, self.halt = None, None, None, False
        self.sm, self.mask, self.sm_max = [], 0, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if self.uuid is not None:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID('GPU-' + str(self.uuid))
                except Exception:
                    handle = None
            if handle is None:
                visible = os.environ.get('CUDA_VISIBLE_DEVICES', '')
                ids = [x for x in visible.split(',') if x.strip().isdigit()]
                phys = int(ids[self.gpu]) if self.gpu < len(ids) else self.gpu
                handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml, self.handle = pynvml, handle
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        while not self.halt:
            try:
                self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                self.mask |= int(self.reasons_fn(self.handle))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.nvml is not None:
            self.halt = True
            self.thread.join(1.0)
            reasons = sorted(nm for bit, nm in self.NVML_REASONS if self.mask & bit)
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.sm_max,
                    'reasons': reasons, 'samples': len(self.sm), 'source': 'nvml, 2 ms period'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi -lms 20'}


# ---- CPU arm: the oracle port of the reference path -------------------------------------------------------
def cpu_reference_arm(steps, warmup, budget_s=20.0):
    """_loss + backward + Adam of the reference network (oracle restatement, torch CPU fp32, all host threads)
    on GUM minibatches of 256 traces.  Returns traces/s over the timed steps."""
    from oracle import network as onet
    from pyprob_b200 import synthetic
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    # parameters with the reference's names/shapes (built on the CPU without the CUDA library)
    torch.manual_seed(0)
    import torch.nn as nn
    params = {}

    def lin(prefix, i, o):
        m = nn.Linear(i, o)
        params[prefix + '.weight'], params[prefix + '.bias'] = m.weight.detach().clone(), m.bias.detach().clone()
    for name in ('obs0', 'obs1'):
        lin('_layers_observe_embedding.{}._layers.0'.format(name), 1, 16)
        lin('_layers_observe_embedding.{}._layers.1'.format(name), 16, 32)
    lin('_layers_observe_embedding_final._layers.0', 64, 64)
    lin('_layers_observe_embedding_final._layers.1', 64, 64)
    I = 64 + 4 + 144
    lstm = nn.LSTM(I, LSTM_DIM, 1)
    for k in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'):
        params['_layers_lstm.' + k] = getattr(lstm, k).detach().clone()
    a = '98__forward__mu__Normal__1'
    params['_layers_address_embedding.' + a] = torch.randn(64)
    params['_layers_distribution_type_embedding.Normal'] = torch.randn(8)
    lin('_layers_sample_embedding.{}._layers.0'.format(a), 1, 4)
    lin('_layers_proposal.{}._ff._layers.0'.format(a), LSTM_DIM, (LSTM_DIM + 30) // 2)
    lin('_layers_proposal.{}._ff._layers.1'.format(a), (LSTM_DIM + 30) // 2, 30)
    plist = {k: v.requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(plist.values()), lr=1e-3)

    def one_step():
        b = synthetic.gum_batch(rng, BATCH)
        subs = [{k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sb.items()} for sb in b.subs]
        opt.zero_grad()
        loss, _ = onet.loss(plist, subs, ['obs0', 'obs1'], [1, 1], 10)
        loss.backward()
        opt.step()
        return float(loss.detach())
    # "all the host threads it can use": pick the fastest intra-op thread count for this (small-GEMM) workload
    best = None
    for nt in sorted(set([1, 2, 4, 8, 16, 32, 64, threads])):
        if nt > threads:
            continue
        torch.set_num_threads(nt)
        one_step()
        dt = float('inf')
        for _ in range(3):   # best of three: a single step is too noisy to choose on
            t0 = time.perf_counter()
            one_step()
            dt = min(dt, time.perf_counter() - t0)
        if best is None or dt < best[1]:
            best = (nt, dt)
    threads = best[0]
    torch.set_num_threads(threads)
    for _ in range(max(warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one_step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {'value': done * BATCH / dt, 'unit': 'traces/s', 'cores': threads, 'kind': 'port',
            'sample': '{} steps of _loss+backward+Adam on {}-trace GUM minibatches (oracle/network.py, torch CPU fp32, '
                      '{} threads), {:.1f} s'.format(done, BATCH, threads, dt)}, dt / done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', type=int, default=0)
    ap.add_argument('--no-extra', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=15.0, help='seconds of CPU work for the cpu_baseline sample')
    ap.add_argument('--nccl-allreduce', action='store_true',
                    help='N>1: NCCL all-reduce + local Adam instead of the fused peer-memory optimiser step')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(args.warmup, 3)

    if args.impl == 'reference':
        if rank != 0:
            return
        cb, s_per_step = cpu_reference_arm(args.steps, warmup, budget_s=120.0)
        print(json.dumps({'impl': 'reference', 'metric': 'ic_train_traces_per_sec', 'value': cb['value'],
                          'unit': 'traces/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': warmup,
                          'ms_per_step': s_per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': WORKLOAD, 'global_batch': BATCH, 'note': 'CPU oracle port of the '
                                     'reference path; rank 0 only'},
                          'cpu_baseline': cb,
                          'e2e': {'value': cb['value'], 'unit': 'traces/s', 'h2d_bytes_per_step': 0,
                                  'd2h_bytes_per_step': 0}}))
        return

    import torch.distributed as dist
    from pyprob_b200 import _lib, synthetic
    from pyprob_b200._lib import call, ptr
    from pyprob_b200.util import Optimizer
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    rng = np.random.default_rng(1234 + rank)
    net = synthetic.gum_network(lstm_dim=LSTM_DIM, precision=args.precision, seed=0)
    net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
    net._create_optimizer()
    net._sync_native()
    if world > 1:
        dist.broadcast(net._arena.data, 0)
    nparams = net._arena.numel()
    peer = None
    if world > 1 and not args.nccl_allreduce:
        # data-parallel step = ONE kernel over NVLink peer memory: reduce-scatter + Adam on the owned slice +
        # all-gather of the parameters (ppb_dp_adam_step); the arena and the gradient live in the peer block
        from pyprob_b200 import parallel
        peer = parallel.PeerAdam(nparams, dev)
        peer.params.copy_(net._arena.data)
        net._arena_store = peer.params
        net._arena = torch.nn.Parameter(peer.params)
        grad = peer.grad[:nparams]
    else:
        grad = torch.zeros(nparams, device=dev)
    net._arena.grad = grad
    batches = [synthetic.gum_batch(rng, BATCH) for _ in range(4)]
    encs = [b.encode(net) for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step -----------------------------------------------------------------------------
    import ctypes as C
    from pyprob_b200.network import BatchStruct
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0 / world], dtype=torch.float32, device=dev)
    adam_state = torch.zeros(4, dtype=torch.int32, device=dev)
    # the loss rides in the gradient tail so that the fused collective sums it with the gradient
    loss = peer.grad[nparams:nparams + 1].view(()) if peer is not None else torch.empty((), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    # all batches of the workload share one structure: one "current batch" image in HBM is refreshed (device to
    # device) from the resident batches, so the index/problem lists are built and uploaded once
    hosts = [torch.from_numpy(enc.pack().copy()).pin_memory() for enc in encs]
    resident = [h.to(dev) for h in hosts]
    cur = torch.empty_like(resident[0])
    bs = BatchStruct()
    call('ppb_batch_from_image', hosts[0].data_ptr(), cur.data_ptr(), hosts[0].numel(), C.byref(bs))
    need = net._ensure_workspace(encs[0])

    def device_step(i):
        cur.copy_(resident[i % 4], non_blocking=True)
        grad.zero_()
        call('ppb_ic_loss_forward', net._handle, ptr(net._arena.data), C.byref(bs), ptr(net._workspace), need,
             args.precision, ptr(loss), ptr(status), None, 1, torch.cuda.current_stream().cuda_stream)
        call('ppb_ic_loss_backward', net._handle, ptr(net._arena.data), ptr(grad), C.byref(bs), ptr(net._workspace), need,
             args.precision, 1.0, torch.cuda.current_stream().cuda_stream)
        if peer is not None:
            peer.step(net._exp_avg, net._exp_avg_sq, hyper, adam_state, torch.cuda.current_stream().cuda_stream)
            return
        if world > 1:
            dist.all_reduce(grad)
        call('ppb_adam_step_dev', ptr(net._arena.data), ptr(grad), ptr(net._exp_avg), ptr(net._exp_avg_sq), nparams,
             ptr(hyper), ptr(adam_state), torch.cuda.current_stream().cuda_stream)

    for i in range(warmup):
        device_step(i)
    barrier()
    use_graph = not args.no_graph
    graphs = []
    if use_graph:  # the whole step (incl. the NCCL all-reduce for N > 1) replays from one CUDA graph per resident batch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(4):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    device_step(i)
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        for i in range(warmup):
            graphs[i % 4].replay()
    barrier()

    def run_step(i):
        if use_graph:
            graphs[i % 4].replay()
        else:
            device_step(i)
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), 'uuid', None))
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the timed spans)
        ev[i][0].record(stream)
        run_step(i)
        ev[i][1].record(stream)
    barrier()
    # graph replays bypass the library's host-side launch counter: count the launches of one eager step
    l0 = _lib.call('ppb_launch_count')
    device_step(0)
    torch.cuda.synchronize()
    launches = (_lib.call('ppb_launch_count') - l0) * args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = args.steps * BATCH * world / (dev_ms * 1e-3)

    # ---- e2e: C-ABI host-buffer call, H2D + D2H inside the timed region --------------------------------------
    loss_host = torch.zeros(1).pin_memory()
    status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    e2e_img_dev = cur
    ws_bytes = net._workspace.numel()
    step_no = [int(adam_state.view(torch.int64)[0].item())]

    def e2e_step(i):
        host = hosts[i % 4]
        if world == 1:
            step_no[0] += 1
            call('ppb_ic_train_step_host', net._handle, ptr(net._arena.data), ptr(grad), ptr(net._exp_avg),
                 ptr(net._exp_avg_sq), nparams, host.data_ptr(), host.numel(), ptr(e2e_img_dev), ptr(net._workspace),
                 ws_bytes, args.precision, 1e-3, 0.9, 0.999, 1e-8, 0.0, step_no[0], loss_host.data_ptr(),
                 status_host.data_ptr(), stream.cuda_stream)
        else:
            resident[i % 4].copy_(host, non_blocking=True)   # host -> device copy of this step's batch image
            device_step(i)
            loss_host.copy_(loss.view(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
    for i in range(warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        e2e_step(i)
    e1.record(stream)
    wall_ms = (time.perf_counter() - t0) * 1e3   # every e2e step ends with a stream synchronize
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)   # what the caller waits for: the slower of device and host clocks
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = args.steps * BATCH * world / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    if peer is not None and peer.timed_out():
        raise RuntimeError('fused data-parallel step: a cross-rank barrier timed out; the measurement is void')

    # ---- roofline of the gate-GEMM class: per-launch durations from CUDA events inside the step ---------------
    peaks = measured_peaks()
    roof = None
    # every rank runs the profiled steps (the step contains the collective); only rank 0 records and reports
    if rank == 0:
        call('ppb_prof_enable', 1)
    prof_steps = min(args.steps, 20)
    for i in range(prof_steps):
        flush.zero_()
        device_step(i)
    torch.cuda.synchronize()
    if rank == 0:
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        call('ppb_prof_read', C.byref(ms), C.byref(n), C.byref(fl))
        call('ppb_prof_enable', 0)
        tf32_peak = peaks['bf16_tflops_sustained'] / 2.0   # tf32 dense = 1/2 of the measured bf16 GEMM peak
        achieved = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
        roof = {'bound': 'tensor', 'kernel': 'LSTM gate GEMM class (P_obs/P_step/recurrent + their dX/dW)',
                'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf32_peak,
                'traffic': None, 'launches': n.value, 'avg_launch_us': ms.value * 1e3 / max(n.value, 1),
                'flops_per_step': fl.value / max(prof_steps, 1),
                'peak_source': '{} bf16_tflops_sustained / 2 (tf32)'.format(peaks['source'])}

    extra = {}
    cpu_baseline = None
    if rank == 0 and world == 1:
        cpu_baseline, _ = cpu_reference_arm(10 ** 6, 2, budget_s=args.cpu_budget)
        if not args.no_extra:
            # secondary workloads: a failure here must not cost the headline line
            for key, fn in (('posterior', lambda: posterior_extras(dev)),
                            ('scoring_hbm_roofline', lambda: scoring_rooflines(dev, peaks)),
                            ('gate_gemm_saturating_4096x2048x512', lambda: gate_gemm_saturating(dev, peaks))):
                try:
                    res = fn()
                    if key == 'posterior':
                        extra.update(res)
                    else:
                        extra[key] = res
                except Exception as exc:   # noqa: BLE001 - reported in the JSON line
                    extra[key + '_error'] = '{}: {}'.format(type(exc).__name__, exc)
            extra['hbm_peak_gbs'] = peaks['hbm_gbs']
    if rank == 0:
        out = {'metric': 'ic_train_traces_per_sec', 'value': value, 'unit': 'traces/s', 'n_gpus': world,
               'steps': args.steps, 'warmup': warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
               'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'lstm_dim': LSTM_DIM, 'trace_length': 1,
                          'parameters': nparams, 'parallelism': 'dp{}'.format(world),
                          'collective': None if world == 1 else ('fused reduce-scatter+Adam+all-gather over NVLink '
                                                                 'peer memory' if peer is not None else 'nccl all-reduce'),
                          'precision': ['3xTF32', 'TF32', 'fp32-simt'][args.precision],
                          'l2': 'flushed between timed steps (256 MiB memset outside the timed spans)',
                          'cuda_graph': bool(use_graph)},
               'e2e': {'value': e2e_value, 'unit': 'traces/s', 'h2d_bytes_per_step': int(hosts[0].numel()),
                       'd2h_bytes_per_step': 8, 'ms_per_step': e2e_ms / args.steps},
               'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu_baseline,
               'extra': extra}
        print(json.dumps(out))
        sys.stdout.flush()
    if world > 1:
        # captured graphs keep NCCL work objects alive: drop them, drain, and leave without tearing the
        # communicator down (destroy_process_group can block on graph-owned resources)
        graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def scoring_rooflines(dev, peaks):
    """HBM roofline of the scoring / sampling / normalisation kernels at a saturating size (2^24 particles,
    per-particle parameters: every operand array is 64 MiB, the working set is far beyond the 126 MB L2).
    achieved = algorithmic bytes per particle x N / CUDA-event time."""
    from pyprob_b200 import ops
    n, K, C = 1 << 24, 10, 8
    g = torch.Generator(device=dev).manual_seed(0)
    v = torch.randn(n, device=dev, generator=g)
    mu = torch.randn(n, device=dev, generator=g)
    sd = torch.rand(n, device=dev, generator=g) + 0.5
    lo = mu - 2.0
    hi = mu + 2.0
    rate = sd * 4
    cnt = torch.poisson(rate, generator=g)
    probs = torch.rand(n, C, device=dev, generator=g) + 0.01
    cat = torch.randint(0, C, (n,), device=dev, generator=g).float()
    m = torch.randn(n, K, device=dev, generator=g)
    s = torch.rand(n, K, device=dev, generator=g) + 0.1
    p = torch.rand(n, K, device=dev, generator=g) + 0.01
    lw = torch.randn(n, device=dev, generator=g) * 5 - 40
    out = torch.empty(n, device=dev)
    cases = [
        ('normal_log_prob', 16, lambda: ops.normal_log_prob(v, mu, sd, lp_out=out)),
        ('uniform_log_prob', 16, lambda: ops.uniform_log_prob(v, lo, hi, lp_out=out)),
        ('poisson_log_prob', 12, lambda: ops.poisson_log_prob(cnt, rate, lp_out=out)),
        ('categorical_log_prob', 4 * C + 8, lambda: ops.categorical_log_prob(cat, probs, lp_out=out)),
        ('mixture_normal_log_prob', (3 * K + 2) * 4, lambda: ops.mixture_normal_log_prob(v, m, s, p, lp_out=out)),
        ('mixture_truncated_normal_log_prob', (3 * K + 4) * 4,
         lambda: ops.mixture_truncated_normal_log_prob(v, m, s, p, lo, hi, lp_out=out)),
        ('normal_sample', 12, lambda: ops.normal_sample(mu, sd, n, 1, 2)),
        ('weights_finalize', 16, lambda: ops.weights_finalize(lw)),
    ]
    res = []
    for name, bytes_per, fn in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_per * n / (ms * 1e-3) / 1e9
        res.append({'kernel': name, 'bytes_per_particle': bytes_per, 'particles': n, 'ms': ms, 'achieved_gbs': gbs,
                    'frac_of_hbm': gbs / peaks['hbm_gbs']})
    return res


def gate_gemm_saturating(dev, peaks):
    """The LSTM gate GEMM shape at a saturating batch (one recurrent step of 4096 traces: [4096,512] x [512,2048]^T)
    through the production tcgen05 kernel: achieved tensor throughput vs the tf32 roofline."""
    from pyprob_b200 import _lib
    from pyprob_b200._lib import call, ptr, stream
    M, N, K = 4096, 2048, 512
    a = torch.randn(M, K, device=dev)
    b = torch.randn(N, K, device=dev)
    c = torch.empty(M, N, device=dev)

    def pack(x):
        nfl = _lib.call('ppb_packed_floats', x.shape[0], x.shape[1])
        hi = torch.empty(nfl, device=dev)
        lo = torch.empty(nfl, device=dev)
        call('ppb_pack_tf32', ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(hi), ptr(lo), stream())
        return hi, lo
    ah, al = pack(a)
    bh, bl = pack(b)
    out = {}
    for prec, name in ((0, '3xTF32'), (1, 'TF32')):
        for _ in range(5):
            call('ppb_gemm_packed', ptr(ah), ptr(al), ptr(bh), ptr(bl), ptr(c), M, N, K, N, None, 0, prec, stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            call('ppb_gemm_packed', ptr(ah), ptr(al), ptr(bh), ptr(bl), ptr(c), M, N, K, N, None, 0, prec, stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        useful = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        issued = useful * (3 if prec == 0 else 1)
        peak = peaks['bf16_tflops'] / 2.0
        out[name] = {'ms': ms, 'useful_tflops': useful, 'issued_tf32_tflops': issued, 'tf32_peak': peak,
                     'frac_issued': issued / peak}
    return out


def posterior_extras(dev):
    """The metric's other half: IS / IC posterior particles/s through the public Model API (N = 1 only)."""
    import math
    import pyprob_b200 as pyprob
    from pyprob_b200 import InferenceEngine, Model
    from pyprob_b200.distributions import Normal

    class GUM(Model):
        def forward(self):
            mu = pyprob.sample(Normal(1, math.sqrt(5)))
            lik = Normal(mu, math.sqrt(2))
            pyprob.observe(lik, name='obs0')
            pyprob.observe(lik, name='obs1')
            return mu
    pyprob.seed(1)
    pyprob.set_verbosity(0)
    m = GUM()
    out = {}
    for n in (65536, 1 << 24):
        for _ in range(3):
            m.posterior_results(n, observe={'obs0': 8, 'obs1': 9})
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            post = m.posterior_results(n, observe={'obs0': 8, 'obs1': 9})
            ess = post.effective_sample_size  # device->host read of the result
        torch.cuda.synchronize()
        out['is_posterior_particles_per_sec_n{}'.format(n)] = reps * n / (time.perf_counter() - t0)
    # north_star's posterior case: GaussianUnknownMean, IC engine (LSTM h=512), 64k particles
    import contextlib
    import io
    from pyprob_b200 import InferenceNetwork
    with contextlib.redirect_stdout(io.StringIO()):
        m.learn_inference_network(num_traces=10 * 256, batch_size=256, inference_network=InferenceNetwork.LSTM,
                                  lstm_dim=512, observe_embeddings={'obs0': {'dim': 32}, 'obs1': {'dim': 32}})
    n = 65536
    eng = InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK
    for _ in range(2):
        m.posterior_results(n, eng, observe={'obs0': 8, 'obs1': 9})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        post = m.posterior_results(n, eng, observe={'obs0': 8, 'obs1': 9})
        _ = post.effective_sample_size
    torch.cuda.synchronize()
    out['ic_posterior_gum_particles_per_sec_n65536'] = reps * n / (time.perf_counter() - t0)
    out.update(ic_posterior_extra())
    out.update(synthetic50_extra(dev))
    return out


def ic_posterior_extra():
    """BASELINE configs[2] shape: GaussianUnknownMeanMarsaglia (stochastic control flow), IC posterior with 64k
    particles through Model.posterior_results (lock-step while_loop; LSTM h=512).  Throughput does not depend on
    how well the proposals are trained, so the network is only trained long enough to create its layers."""
    import math
    import pyprob_b200 as pyprob
    from pyprob_b200 import InferenceEngine, InferenceNetwork, Model
    from pyprob_b200.distributions import Normal, Uniform

    class Marsaglia(Model):
        def forward(self):
            def body(s):
                x = pyprob.sample(Uniform(-1, 1))
                y = pyprob.sample(Uniform(-1, 1))
                return {'x': x, 'y': y, 's': x * x + y * y}
            st = pyprob.while_loop(lambda s: s['s'] >= 1, body, {'x': 0.0, 'y': 0.0, 's': 2.0})
            mu = 1 + math.sqrt(5) * (st['x'] * torch.sqrt(-2 * torch.log(st['s']) / st['s']))
            lik = Normal(mu, math.sqrt(2))
            pyprob.observe(lik, name='obs0')
            pyprob.observe(lik, name='obs1')
            return mu
    pyprob.seed(3)
    pyprob.set_verbosity(0)
    m = Marsaglia()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m.learn_inference_network(num_traces=20 * 1024, batch_size=1024, inference_network=InferenceNetwork.LSTM,
                                  lstm_dim=512, observe_embeddings={'obs0': {'dim': 32}, 'obs1': {'dim': 32}})
    n = 65536
    m.posterior_results(n, InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK, observe={'obs0': 8, 'obs1': 9})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        post = m.posterior_results(n, InferenceEngine.IMPORTANCE_SAMPLING_WITH_INFERENCE_NETWORK,
                                   observe={'obs0': 8, 'obs1': 9})
        _ = post.effective_sample_size
    torch.cuda.synchronize()
    return {'ic_posterior_marsaglia_particles_per_sec_n65536': reps * n / (time.perf_counter() - t0),
            'ic_posterior_marsaglia_addresses': len(m._inference_network._addresses)}


def synthetic50_extra(dev, B=512, T=50):
    """BASELINE configs[3] shape on one GPU: 50-address Normal/Categorical(4) model, LSTM h=512, obs dim 256,
    512 traces per GPU (the per-GPU share of the 4096-trace global batch on 8 GPUs): device-resident train step."""
    import ctypes as C
    from pyprob_b200 import synthetic
    from pyprob_b200._lib import call, ptr
    from pyprob_b200.network import BatchStruct
    from pyprob_b200.util import Optimizer
    rng = np.random.default_rng(5)
    net = synthetic.synthetic50_network(precision=0, T=T)
    net._optimizer_type, net._learning_rate_init, net._weight_decay = Optimizer.ADAM, 1e-3, 0.0
    net._create_optimizer()
    net._sync_native()
    enc = synthetic.synthetic50_batch(rng, B, T=T).encode(net)
    grad = torch.zeros_like(net._arena.data)
    img = torch.from_numpy(enc.pack().copy()).pin_memory()
    dimg = img.to(dev)
    bs = BatchStruct()
    call('ppb_batch_from_image', img.data_ptr(), dimg.data_ptr(), img.numel(), C.byref(bs))
    need = net._ensure_workspace(enc)
    loss = torch.empty((), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0], dtype=torch.float32, device=dev)
    state = torch.zeros(4, dtype=torch.int32, device=dev)

    def step():
        st = torch.cuda.current_stream().cuda_stream
        grad.zero_()
        call('ppb_ic_loss_forward', net._handle, ptr(net._arena.data), C.byref(bs), ptr(net._workspace), need, 0, ptr(loss),
             ptr(status), None, 1, st)
        call('ppb_ic_loss_backward', net._handle, ptr(net._arena.data), ptr(grad), C.byref(bs), ptr(net._workspace), need, 0,
             1.0, st)
        call('ppb_adam_step_dev', ptr(net._arena.data), ptr(grad), ptr(net._exp_avg), ptr(net._exp_avg_sq),
             net._arena.numel(), ptr(hyper), ptr(state), st)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {'synthetic50_train_traces_per_sec_b512': B / (ms * 1e-3), 'synthetic50_ms_per_step_b512': ms,
            'synthetic50_parameters': int(net._arena.numel())}


if __name__ == '__main__':
    main()
