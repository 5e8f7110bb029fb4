#!/usr/bin/env python
"""Benchmark of the shadowing segment loop (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): Lorenz-96, F = 8, RK4 dt = 0.01,
N = 2^24 sites PER GPU (weak scaling, state sharded along N with halo exchange),
m = 32 homogeneous tangents -> 34 trajectories, 100 time steps per segment,
epsilon = 1e-6 sqrt(N), synthetic data (u0 = RandomState(0).rand(N) + run-up).

One bench "step" = one shadowing segment: the boundary (time dilation, projection,
CholeskyQR2, next perturbed ensemble) + 100 ensemble time steps + the objective
sums / sensitivities brought back to the host.  metric = state-updates/s =
N_total * (m+1) * steps_per_segment * K / time.

`value`   device-resident loop (inputs in HBM when the clock starts).
`e2e`     the public API call continue_shadowing(run, s, checkpoint, ...) on HOST
          numpy buffers: u0/V/v uploaded, K segments, u0/V/v + histories downloaded.
`roofline` the ensemble step kernel: algorithmic bytes 16 B * N_local * (m+2) per
          launch / mean launch duration (CUDA events around every launch in the
          timed region) vs the measured HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline` / --impl reference: the numpy restatement of the reference's loop
          (oracle/shadowing_port.py, pinned against /root/reference by
          oracle/gen_golden.py) on the host cores, process pool of os.cpu_count()
          workers like the reference (fds/segment.py:37), on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'L96 shadowing state-updates/s (N*(m+1)*steps*K / time)'
UNIT = 'state-updates/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5, help='timed segments')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--logn', type=int, default=24, help='log2 of sites per GPU')
    ap.add_argument('--m', type=int, default=32)
    ap.add_argument('--seg-steps', type=int, default=100)
    ap.add_argument('--math', default='exact', choices=['exact', 'fast'])
    ap.add_argument('--cpu-logn', type=int, default=20, help='log2 N of the CPU-baseline sample')
    ap.add_argument('--cpu-seg-steps', type=int, default=10)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-sweep', action='store_true', help='skip the batched-sweep side measurement')
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# CPU leg: the oracle port (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------- #
def cpu_port_run(logn, m, seg_steps, K):
    from oracle import plugins, shadowing_port as port
    N = 1 << logn
    cores = os.cpu_count() or 1
    u0 = np.random.RandomState(0).rand(N)
    u0 = plugins.run_l96_rk4(u0, 0.0, 20)[0]
    np.random.seed(0)
    t0 = time.perf_counter()
    port.shadowing(plugins.run_l96_rk4, u0, 0.0, m, K, seg_steps, 0,
                   epsilon=1e-6 * np.sqrt(N), workers=cores)
    dt = time.perf_counter() - t0
    return N * (m + 1) * seg_steps * K / dt, dt, cores


def cpu_baseline(args):
    K = 2
    v, dt, cores = cpu_port_run(args.cpu_logn, args.m, args.cpu_seg_steps, K)
    return {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/shadowing_port.py shadowing(), L96-RK4 N=2^%d m=%d, %d segments x %d steps, '
                      'pool of %d processes, %.1f s' % (args.cpu_logn, args.m, K, args.cpu_seg_steps, cores, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (numpy port; the Python reference
    itself cannot travel to the GPU box) on the host cores, same metric."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    K = 2
    vals, times = [], []
    for it in range(args.warmup + args.steps):
        v, dt, cores = cpu_port_run(args.cpu_logn, args.m, args.cpu_seg_steps, K)
        if it >= args.warmup:
            vals.append(v), times.append(dt)
    value = float(np.mean(vals))
    sample = ('bounded sample of the workload: N=2^%d (not 2^%d) m=%d, %d segments x %d steps per bench step, '
              'pool of %d processes' % (args.cpu_logn, args.logn, args.m, K, args.cpu_seg_steps, cores))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(times) * 1e3),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def workload_config(args):
    return {'workload': 'Lorenz-96 F=8 RK4 dt=0.01 shadowing, N=2^%d sites per GPU, m=%d (%d trajectories), '
                        '%d steps/segment, eps=1e-6*sqrt(N)' % (args.logn, args.m, args.m + 2, args.seg_steps),
            'N_per_gpu': 1 << args.logn, 'm': args.m, 'steps_per_segment': args.seg_steps, 'math': args.math,
            'bench_step': 'one segment = boundary (time dilation + projection + CholeskyQR2 + next ensemble) '
                          '+ %d ensemble time steps' % args.seg_steps,
            'l2': 'ensemble %.1f GB per GPU >> 126 MB L2 (no flush needed)'
                  % ((1 << args.logn) * (args.m + 2) * 8 / 1e9),
            'parallelism': 'N-sharded ring, %d GPU(s)' % args.gpus}


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.power, self.power_limit = [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                pass
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'power_w': float(np.median(self.power)) if self.power else None, 'power_limit_w': self.power_limit}


# --------------------------------------------------------------------------- #
def djds_check(device):
    """Third part of the BASELINE metric: dJ/ds of config 3 (Lorenz-96 N=40, m=16, 100 segments)
    on the GPU next to the reference's value and its own error bar (committed fixture generated
    from the reference by oracle/gen_golden.py); pass = ratio <= 1."""
    import fds_b200
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'cfg3_l96rk4_N40_m16_K100.npz'))
    np.random.seed(int(g['seed']))
    run = fds_b200.L96(40, device=device, distributed=False)
    J, G = fds_b200.shadowing(run, g['u0'], float(g['s']), int(g['m']), int(g['K']), int(g['steps']),
                              int(g['runup']), epsilon=float(g['eps']))
    run.close()
    err = np.abs(G - g['G_ref'])
    return {'config': 'Lorenz-96 N=40 F=8.1 RK4, m=16, 100 segments x 100 steps (BASELINE configs[2])',
            'G': G.tolist(), 'G_ref': g['G_ref'].tolist(), 'sigma_ref': g['G_err'].tolist(),
            'abs_err': err.tolist(), 'ratio': (err / g['G_err']).tolist(),
            'pass': bool((err <= g['G_err']).all())}


def sweep_check(device, B=4096, N=4096, m=16, K=2, steps=100):
    """BASELINE configs[4] measured beside the headline: B independent Lorenz-96 problems (F swept
    over [6, 10]) in one batched context; whole segments (boundary + 100 steps), device resident."""
    import fds_b200
    eps = 1e-4
    run = fds_b200.L96(N, device=device, distributed=False)
    eng = run.batch_engine(m, B)
    eng.ctx.set_batch_params(np.linspace(-2.0, 2.0, B))
    eng.set_state(np.random.RandomState(0).rand(B, N))
    eng.run(0.0, 200, want_J=False)
    eng.orthonormal_random_tangents(seed=1)
    eng.boundary(0.0, eps, 0, 0, 1)
    eng.segment(0.0, eps, steps)

    def seg():
        R = eng.boundary(0.0, eps, 1, 1, 1, keep_tangents=False)[0]
        return R, eng.segment_sens(0.0, eps, steps)[1]
    seg()
    eng.ctx.sync()
    t0 = time.perf_counter()
    for _ in range(K):
        R, J = seg()
    eng.ctx.sync()
    dt = time.perf_counter() - t0
    ok = bool(np.isfinite(J).all() and np.isfinite(R).all())
    run.close()
    return {'config': '%d independent Lorenz-96 N=%d problems, m=%d, one batched context on one GPU' % (B, N, m),
            'value': B * N * (m + 1) * steps * K / dt, 'unit': UNIT, 'segments': K, 'seconds': dt, 'finite': ok}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import fds_b200
    from fds_b200.segment import sensitivities

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one process per GPU on a multi-socket host: run on the CPUs next to this rank's GPU, so that the
        # page-locked host buffers of the end-to-end leg (allocated below) sit on the GPU's own NUMA node and
        # eight 4.6 GB uploads do not all cross the socket interconnect.  (Not at N = 1: the cpu_baseline leg
        # that follows uses every core.)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception as e:      # affinity is an optimisation, never a requirement
            print('bench: no NUMA affinity (%s)' % e, file=sys.stderr)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus)

    n_local = 1 << args.logn
    N = n_local * world
    m, M, S = args.m, args.m + 2, args.seg_steps
    eps = 1e-6 * np.sqrt(N)
    s = 0.0
    stream = torch.cuda.current_stream().cuda_stream
    run = fds_b200.L96(N, math=args.math, device=local_rank, stream=stream)
    eng = run.engine(m)
    lo, hi = eng.lo, eng.hi

    # synthetic initial state on the attractor; random orthonormal tangents (device RNG)
    eng.ctx.set_state(np.random.RandomState(rank).rand(n_local))
    eng.run(s, 200, want_J=False)
    eng.orthonormal_random_tangents(seed=1)
    eng.boundary(s, eps, from_ensemble=0, do_qr=0, build=1)
    eng.segment(s, eps, S)

    def one_segment():
        R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=1, do_qr=1, build=1, keep_tangents=False)
        J0, G, g = sensitivities(eng.segment(s, eps, S), eps)
        return R, b, Gd, gd, J0, G, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_segment()
    sampler = ClockSampler(local_rank)
    eng.ctx.profile_enable(True)
    l0 = eng.ctx.launches
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hist = [one_segment() for _ in range(args.steps)]
    e1.record()
    barrier()
    clocks = sampler.result()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
    launches = torch.tensor([eng.ctx.launches - l0], device='cuda', dtype=torch.float64)
    kern_ms, kern_n = eng.ctx.profile_read()
    steps_per_launch = eng.ctx.profile_steps / max(kern_n, 1)     # 2 with temporal blocking
    eng.ctx.profile_enable(False)
    kern = torch.tensor([kern_ms / max(kern_n, 1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = N * (m + 1) * S * args.steps / (total_ms * 1e-3)

    # sanity of what was computed inside the timed region
    R = hist[-1][0]
    assert np.isfinite(R).all() and (np.diag(R) > 0).all()
    assert np.isfinite(hist[-1][4]).all()

    # ---- end to end through the public API on host buffers ----------------- #
    e2e = per_seg = None
    if not args.no_e2e:
        from fds_b200.lsstan import LssTangent
        from fds_b200.checkpoint import Checkpoint
        # host inputs in page-locked memory (get_* return pinned arrays; global arrays would be
        # sliced per rank inside the API)
        V, v = eng.ctx.get_tangents()
        u = eng.ctx.get_state()
        plug = _LocalShardPlugin(run, eng)

        def call():
            cp = Checkpoint(u, V, v, LssTangent(), [], [], [], [], [])
            return fds_b200.continue_shadowing(plug, s, cp, args.steps, S, epsilon=eps, return_checkpoint=True)

        out = call()          # untimed: first use allocates the pinned result buffers (recycled afterwards)
        del out
        barrier()
        t0 = time.perf_counter()
        out = call()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        assert out.lss.K_segments() == args.steps and np.isfinite(out.V).all()
        # strictest reading: the reference's run_segment(u0, V, v) -> (u0, V, v, J0, G, g) on host arrays, i.e.
        # the whole ensemble state crosses PCIe in both directions EVERY segment (single GPU only)
        bytes_in_1 = M * n_local * 8
        if world == 1:
            from fds_b200.segment import run_segment
            run_segment(run, u, V, v, s, 0, S, eps)
            barrier()
            t1 = time.perf_counter()
            res = None
            for _ in range(2):
                res = None                 # hand the pinned result buffers back to the pool before the next call
                res = run_segment(run, u, V, v, s, 0, S, eps)
            barrier()
            dt1 = (time.perf_counter() - t1) / 2
            assert np.isfinite(res[1]).all()
            per_seg = {'value': N * (m + 1) * S / dt1, 'unit': UNIT, 'seconds_per_segment': dt1,
                       'h2d_bytes_per_step': float(bytes_in_1), 'd2h_bytes_per_step': float(bytes_in_1 + S * M * 16),
                       'call': 'segment.run_segment(run, u0, V, v, s, 0, %d, eps) on pinned host arrays, one call per segment' % S}
        bytes_in = (M * n_local * 8) * world
        small = args.steps * (m * m + m + m + 1 + S * M * 2) * 8
        e2e = {'value': N * (m + 1) * S * args.steps / dt, 'unit': UNIT,
               'h2d_bytes_per_step': bytes_in / args.steps, 'd2h_bytes_per_step': (bytes_in + small) / args.steps,
               'call': 'continue_shadowing(run, s, checkpoint(host u0,V,v), %d segments, return_checkpoint=True); '
                       'one call = %d bench steps, %.3f s' % (args.steps, args.steps, dt)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        kms = float(kern.item())
        alg_bytes = 16.0 * n_local * M * steps_per_launch
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        traffic = fp64_pct = None
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')))
            traffic, fp64_pct = prof['dram_bytes_per_launch'], prof.get('fp64_pipe_busy_pct')
        except Exception:
            pass
        # FP64-pipe utilisation of the step kernel from the live numbers: FP64 warp-instructions per launch
        # (static count of the hot loop, DESIGN 4.1: 476 per 8-site block of two steps = 29.75 per update in exact
        # mode with the objective sums) over what 4 schedulers x 1 instruction per 2 clocks issue at the SM clock
        # sampled during the timed region
        fp64_live = None
        if clocks.get('sm_mhz') and args.math == 'exact' and steps_per_launch == 2 and torch.cuda.is_available():
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            winst = n_local * M * steps_per_launch * 29.75 / 32.0
            fp64_live = 100.0 * winst / (kms * 1e-3 * clocks['sm_mhz'] * 1e6 * sms * 4 * 0.5)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(args),
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                             'frac': achieved / peak, 'traffic': traffic,
                             'kernel': 'l96_rk4_stream_kernel', 'kernel_ms': kms,
                             'algorithmic_bytes_per_launch': alg_bytes, 'time_steps_per_launch': steps_per_launch,
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if peaks else 'fallback 6650 GB/s',
                             'frac_of_nominal_8TBs': achieved / 8000.0,
                             'fp64_pipe_busy_pct_ncu': fp64_pct, 'fp64_pipe_busy_pct_live': fp64_live,
                             'step_kernel_share_of_time': kms * (S / steps_per_launch) * args.steps / total_ms,
                             'note': ('temporal blocking: one launch advances %g time steps, so DRAM traffic per launch '
                                      '(`traffic`) is 1/%g of the algorithmic bytes and `frac` can exceed 1; the kernel is '
                                      'then bound by the FP64 pipe (profiles/)' % (steps_per_launch, steps_per_launch))
                                     if steps_per_launch > 1 else 'one time step per launch'},
                'clocks': clocks, 'gpu_launches': int(launches.item())}
        if e2e is not None:
            line['e2e'] = e2e
            if per_seg is not None:
                line['e2e_run_segment'] = per_seg
        line['djds'] = djds_check(local_rank)
        if world == 1 and not args.no_sweep:
            line['sweep_cfg5'] = sweep_check(local_rank)
        if not args.no_cpu and world == 1:
            line['cpu_baseline'] = cpu_baseline(args)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _LocalShardPlugin:
    """Hands the already-built engine to continue_shadowing (per-rank shard arrays in, out)."""

    def __init__(self, run, eng):
        self.run, self.eng = run, eng

    def engine(self, m):
        assert m == self.eng.m
        return _LocalEngine(self.eng)


class _LocalEngine:
    """Engine view whose host arrays are per-rank shards (no global gather of 4.6 GB blocks)."""

    def __init__(self, eng):
        self.__dict__['e'] = eng

    def __getattr__(self, k):
        return getattr(self.e, k)

    def get_state(self):
        return self.e.ctx.get_state()

    def get_tangents(self):
        return self.e.ctx.get_tangents()

    def set_state(self, u):
        self.e.ctx.set_state(u)

    def set_tangents(self, V, v):
        self.e.ctx.set_tangents(V, v)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints at the C level
    (e.g. NCCL's version banner) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1

if __name__ == '__main__':
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
