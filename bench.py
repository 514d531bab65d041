#!/usr/bin/env python
"""
bench.py -- headline benchmark of the hot path (BASELINE.json: "NUFFT fwd+adj pairs/s at 3D 128^3 J=6;
interp/gridding HBM GB/s vs peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one forward + one adjoint NUFFT (a "pair") on configuration 3 of BASELINE.json:
Nd 128^3, Kd 256^3, Jd 6^3, M = 2,000,000 uniform random samples, complex64, one coil per GPU.
At N > 1 (torchrun, one rank per GPU, NCCL) every rank owns one coil of the same trajectory (weak
scaling: per-GPU work fixed) and the adjoint image is summed with one all-reduce per step, which is
the only exchange the path has (adjoint_many2one).  Prints ONE JSON line on rank 0.

  value      device-resident pairs/s over all ranks: EXACTLY --steps pairs, CUDA events, max over ranks
  sustained  the same loop repeated for >= 1 s (clocks are sampled over both regions)
  e2e        the same through the host API (pinned host buffers, H2D + D2H of every step inside the timed region);
             N > 1: forward_one2many / adjoint_many2one semantics -- ONE host image in (H2D on rank 0, NCCL broadcast),
             per-coil data out and in on every rank, ONE reduced image out (D2H on rank 0)
  roofline   dominant kernel (slower of interp / gridding): algorithmic bytes (SURVEY.md 8d:
             8*K + 12*M*sum(J) + 8*M = 582.2 MB) / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's numpy/scipy CPU path on a bounded sample
  config5_cg / config2  extra keys: configuration-5 CG (32 coils sharded over the N ranks) ms per iteration,
             configuration-2 stage times (N = 1)
--impl reference times that CPU path alone on the FULL workload (rank 0 only): every step evaluates all 2 M samples
(interpolation and gridding in chunks of 500 k rows; plan excluded), nothing is extrapolated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ND, KD, JD, M = (128, 128, 128), (256, 256, 256), (6, 6, 6), 2_000_000
WORKLOAD = '3D Nd=128^3 Kd=256^3 Jd=6^3 M=2000000 uniform-random samples, 1 coil per GPU (BASELINE.json configs[2])'
ALGO_BYTES = 8 * 256 ** 3 + 12 * M * 18 + 8 * M          # interp or gridding, SURVEY.md 8d
CPU_SAMPLE_M = 100_000
CPU_CHUNK = 500_000


def make_config(n_gpus):
    """the SAME dict in both arms (ours / reference)"""
    return {'workload': WORKLOAD,
            'l2_policy': 'working set per step (134 MB grid + 192 MB + 256 MB plan records) exceeds the 126 MB L2',
            'parallelism': ('coil-sharded x%d, one all-reduce of the adjoint image per step' % n_gpus) if n_gpus > 1
            else 'single GPU'}


def make_om():
    return numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference's numpy/scipy implementation)
# ------------------------------------------------------------------------------------------------
class CpuPath:
    """Full-size pad/FFT/crop (256^3) + CSR interpolation / gridding on the first `m` of the 2 M samples, evaluated in
    chunks of <= 500 k rows (forward rows are independent, the adjoint is the sum of the per-chunk spH . y;
    BASELINE.md section 3).  m = M: the whole workload, nothing scaled.  m < M (the bounded cpu_baseline leg of the GPU
    arm): the interpolation / gridding time is scaled by M / m (CSR SpMV cost is linear in the rows).
    threads > 1: the chunks run on a thread pool (scipy's CSR kernels release the GIL), each thread summing its own
    chunks' adjoint grids before the partial grids are added; the FFTs stay numpy.fft as in the reference."""

    def __init__(self, m, threads=1):
        from concurrent.futures import ThreadPoolExecutor
        from oracle import nufft_oracle as orc
        om = make_om()[:m]
        self.m = m
        self.threads = max(1, threads)
        nchunks = max((m + CPU_CHUNK - 1) // CPU_CHUNK, self.threads if m == M else 1)
        rows = (m + nchunks - 1) // nchunks
        self.rows = rows

        def plan(s):
            O = orc.NUFFT()
            O.plan(om[s:s + rows], ND, KD, JD)
            return O
        self.pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None
        starts = list(range(0, m, rows))
        self.parts = list(self.pool.map(plan, starts)) if self.pool else [plan(s) for s in starts]
        rng = numpy.random.default_rng(1)
        self.x = (rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)
        self.scale = M / m

    def pair_seconds(self):
        O0 = self.parts[0]
        t = time.perf_counter
        t0 = t()
        k = O0.xx2k(O0.x2xx(self.x))
        t1 = t()
        if self.pool:
            ys = list(self.pool.map(lambda O: O.k2y(k), self.parts))
        else:
            ys = [O.k2y(k) for O in self.parts]
        t2 = t()

        def grid_some(idx):                 # one thread: its chunks, summed as they come
            acc = None
            for i in idx:
                kk = self.parts[i].y2k(ys[i].astype(numpy.complex64))
                acc = kk if acc is None else acc + kk
            return acc
        if self.pool:
            groups = [list(range(j, len(self.parts), self.threads)) for j in range(min(self.threads, len(self.parts)))]
            partial = list(self.pool.map(grid_some, groups))
            k2 = partial[0]
            for kk in partial[1:]:
                k2 = k2 + kk
        else:
            k2 = grid_some(range(len(self.parts)))
        t3 = t()
        O0.xx2x(O0.k2xx(k2))
        t4 = t()
        return (t1 - t0) + (t4 - t3) + self.scale * ((t2 - t1) + (t3 - t2))

    def describe(self):
        if self.m == M:
            return ('oracle port of the reference numpy/scipy CPU path, FULL workload: scale+pad+fftn(256^3)+ifftn+crop and '
                    'CSR interpolation+gridding of all %d samples in %d chunks of <= %d rows (plan excluded); numpy.fft as the '
                    'reference (one thread), scipy CSR SpMV chunks on %d thread(s)' % (M, len(self.parts), self.rows, self.threads))
        return ('oracle port of the reference numpy/scipy CPU path: full-size scale+pad+fftn(256^3)+ifftn+crop, '
                'CSR interpolation+gridding on %d of %d samples scaled x%d; single-threaded like the reference '
                '(numpy.fft + scipy CSR SpMV)' % (self.m, M, M // self.m))


def run_reference(args, rank):
    if rank != 0:
        return
    t_plan = time.perf_counter()
    threads = max(1, min(os.cpu_count() or 1, 16))
    cpu = CpuPath(M, threads)
    t_plan = time.perf_counter() - t_plan
    for _ in range(max(args.warmup, 0)):
        cpu.pair_seconds()
    ts = [cpu.pair_seconds() for _ in range(args.steps)]
    sec = float(numpy.mean(ts))
    v = 1.0 / sec
    line = {
        'impl': 'reference', 'metric': 'NUFFT forward+adjoint pairs/s', 'value': v, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
        'config': make_config(args.gpus),
        'cpu_baseline': {'value': v, 'unit': 'pairs/s', 'cores': threads, 'cores_available': os.cpu_count(),
                         'kind': 'port', 'sample': cpu.describe()},
        'e2e': {'value': v, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'plan_seconds': t_plan, 'measured': 'every step evaluated in full; nothing extrapolated',
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [v.strip() for v in r.split(',')]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except Exception:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import pynufft_b200
    from pynufft_b200 import _lib

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    om = make_om()
    A = pynufft_b200.NUFFT(dev)
    t0 = time.perf_counter()
    A.plan(om, ND, KD, JD)
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0

    rng = numpy.random.default_rng(100 + rank)
    x_host = torch.from_numpy((rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)).pin_memory()
    x = x_host.to(dev)

    pending = [None]

    def step():
        y = A._forward_device(x)
        xa = A._adjoint_device(y)
        if dist is not None:
            # adjoint_many2one image sum over the coil shards: asynchronous on NCCL's stream, so that it overlaps the
            # next step's kernels (nothing in the next step reads it); at most one reduction is in flight
            if pending[0] is not None:
                pending[0].wait()
            pending[0] = dist.all_reduce(xa, async_op=True)
        return xa

    def drain():
        if pending[0] is not None:
            pending[0].wait()              # the current stream waits for the last reduction: it is inside the timed region
            pending[0] = None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, iters, warm, fin=None):
        for _ in range(warm):
            fn()
        if fin is not None:
            fin()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        if fin is not None:
            fin()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- headline: device-resident pairs ----
    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    clocks = ClockSampler(local_rank)
    time.sleep(0.25)
    l0 = lib.b200nufft_launch_count()
    tw0 = time.perf_counter()
    ms = timed(step, args.steps, 0, drain)
    launches = lib.b200nufft_launch_count() - l0
    value = world * args.steps / (ms * 1e-3)
    # the same loop sustained for >= 1 s (the K-step region above is only K * 0.7 ms long)
    sus_iters = max(args.steps, int(1.2 / max(ms / args.steps * 1e-3, 1e-6)))
    ms_sus = timed(step, sus_iters, 0, drain)
    tw1 = time.perf_counter()
    clk = clocks.stop(tw0, tw1)
    sustained = {'value': world * sus_iters / (ms_sus * 1e-3), 'unit': 'pairs/s', 'steps': sus_iters,
                 'seconds': ms_sus * 1e-3, 'ms_per_step': ms_sus / sus_iters}

    # ---- e2e: host API, pinned host buffers, H2D + D2H of every step inside the timed region ----
    # A step = forward of a host image + adjoint of a host data vector, each delivered back to the host.  NS steps are in
    # flight, so the copy-in, the kernels and the copy-out of neighbouring steps overlap; every step still moves all of
    # its inputs and outputs over PCIe.
    def pinned(shape):
        return torch.empty(shape, dtype=torch.complex64).pin_memory()
    NS = 3                                                       # steps in flight (staging slots used)
    x_hosts = [x_host] + [pinned(ND) for _ in range(NS - 1)]
    y_in = [pinned((M,)) for _ in range(NS)]
    y_out = [pinned((M,)) for _ in range(NS)]
    xa_out = [pinned(ND) for _ in range(NS)]
    A.forward(x_hosts[0].numpy(), out=y_in[0].numpy())          # realistic adjoint input: a forward result
    for s_ in range(1, NS):
        x_hosts[s_].copy_(x_host)
        y_in[s_].copy_(y_in[0])
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    if dist is None:
        def e2e_run(iters):
            for i in range(iters):
                s = i % NS
                if i >= NS:
                    A.wait('forward', s)
                    A.wait('adjoint', s)
                A.forward(x_hosts[s].numpy(), out=y_out[s].numpy(), slot=s)
                A.adjoint(y_in[s].numpy(), out=xa_out[s].numpy(), slot=s)
            for s in range(NS):
                A.wait('forward', s)
                A.wait('adjoint', s)
        h2d = x_host.numel() * 8 + y_in[0].numel() * 8
        d2h = y_out[0].numel() * 8 + xa_out[0].numel() * 8
        e2e_mode = 'pipelined host API (NUFFT.forward / adjoint(..., slot) + wait), %d steps in flight' % NS
    else:
        # coil-sharded one2many / many2one through the host boundary: the ONE image enters on rank 0 and is broadcast
        # over NVLink; every rank returns its coil's data (D2H) and takes its coil's data (H2D); the all-reduced image
        # leaves from rank 0 only.  Per-rank PCIe traffic: rank 0 65.6 MB, the others 32 MB (the per-coil data).
        x_devs = [torch.empty(ND, dtype=torch.complex64, device=dev) for _ in range(NS)]
        y_devs = [torch.empty((M,), dtype=torch.complex64, device=dev) for _ in range(NS)]
        ev_in = [torch.cuda.Event() for _ in range(NS)]
        ev_yin = [torch.cuda.Event() for _ in range(NS)]
        ev_comp = [torch.cuda.Event() for _ in range(NS)]
        ev_fwd = [torch.cuda.Event() for _ in range(NS)]
        ev_out = [torch.cuda.Event() for _ in range(NS)]
        keep = [None] * NS

        def dist_step(s):
            cur = torch.cuda.current_stream()
            s_in.wait_event(ev_comp[s])                      # the operators that last read slot s are done
            with torch.cuda.stream(s_in):
                if rank == 0:
                    x_devs[s].copy_(x_hosts[s], non_blocking=True)
                ev_in[s].record()
                y_devs[s].copy_(y_in[s], non_blocking=True)
                ev_yin[s].record()
            cur.wait_event(ev_in[s])
            dist.broadcast(x_devs[s], src=0)                 # NCCL over NVLink instead of N host copies of the image
            y = A._forward_device(x_devs[s])
            ev_fwd[s].record()
            cur.wait_event(ev_yin[s])
            xa = A._adjoint_device(y_devs[s])
            ev_comp[s].record()
            work = dist.all_reduce(xa, async_op=True)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_fwd[s])
                y_out[s].copy_(y, non_blocking=True)
                work.wait()                                  # the copy-out stream (not the compute stream) waits for NCCL
                if rank == 0:
                    xa_out[s].copy_(xa, non_blocking=True)
                ev_out[s].record()
            y.record_stream(s_out)
            xa.record_stream(s_out)
            keep[s] = (y, xa)

        def e2e_run(iters):
            for i in range(iters):
                s = i % NS
                if i >= NS:
                    ev_out[s].synchronize()
                dist_step(s)
            for s in range(NS):
                ev_out[s].synchronize()
        # bytes per step summed over the ranks
        h2d = x_host.numel() * 8 + world * y_in[0].numel() * 8
        d2h = world * y_out[0].numel() * 8 + xa_out[0].numel() * 8
        e2e_mode = ('coil-sharded forward_one2many + adjoint_many2one through the host boundary, %d steps in flight: image '
                    'H2D on rank 0 + NCCL broadcast, per-coil data D2H / H2D on every rank, all-reduce, image D2H on '
                    'rank 0; bytes are summed over the ranks' % NS)
    e2e_iters = max(2 * NS, min(args.steps, 60))
    e2e_run(2 * NS)
    ms_e2e = timed(lambda: e2e_run(e2e_iters), 1, 0)
    e2e_value = world * e2e_iters / (ms_e2e * 1e-3)

    # the blocking host calls (one step at a time, no overlap), for comparison
    def e2e_blocking():
        A.forward(x_hosts[0].numpy(), out=y_out[0].numpy())
        A.adjoint(y_in[0].numpy(), out=xa_out[0].numpy())
    blk_iters = 10
    ms_blk = timed(e2e_blocking, blk_iters, 2) if dist is None else None

    # ---- per-kernel times (rank-local, CUDA events on the launch stream) ----
    P = ctypes_ptr
    st = lambda: P(torch.cuda.current_stream().cuda_stream)
    grid = torch.empty((1,) + KD, dtype=torch.complex64, device=dev)
    yv = torch.empty((M,), dtype=torch.complex64, device=dev)
    xo = torch.empty(ND, dtype=torch.complex64, device=dev)
    lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st())
    kit = max(10, min(args.steps, 50))
    kern = {}
    kern['cufft_3d_full'] = timed(lambda: lib.b200nufft_fft(A._plan, P(grid.data_ptr()), 1, 0, st()), kit, 3) / kit
    # scale + pad + pruned FFT (three fused passes at Kd=256^3); leaves a well-scaled grid for the interp timing
    kern['pad_fft_true_grid'] = timed(lambda: lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st()), kit, 3) / kit
    kern['interp_true_grid'] = timed(lambda: lib.b200nufft_interp(A._plan, P(grid.data_ptr()), P(yv.data_ptr()), 1, st()), kit, 3) / kit
    # the forward's own pair of stages: the fused passes emit the phase-modulated grid the column-sweep gather reads
    # (csrc/col3d.cu); "pad_fft" / "interp" = the stages as `forward` runs them, *_true_grid = the x2xx+xx2k / k2y stages
    # of the API on a plain grid (tiled gather, csrc/interp_tiled.cu)
    if A._kspace_modulated():
        gm = torch.empty((1,) + KD, dtype=torch.complex64, device=dev)
        kern['pad_fft'] = timed(lambda: lib.b200nufft_pad_fft_modulated(A._plan, P(x.data_ptr()), P(gm.data_ptr()), 1, 1, 0, None, st()), kit, 3) / kit
        kern['interp'] = timed(lambda: lib.b200nufft_interp_modulated(A._plan, P(gm.data_ptr()), P(yv.data_ptr()), 1, st()), kit, 3) / kit
        del gm
    else:
        kern['pad_fft'], kern['interp'] = kern['pad_fft_true_grid'], kern['interp_true_grid']
    # the adjoint's own pair of stages: the column-sweep gridding leaves the grid phase-modulated and the fused inverse
    # passes undo it (csrc/col3d.cu); "gridding" = the whole stage as `adjoint` runs it: pre-pass (grid zero-fill, sorted
    # data gather) + scatter kernel.  gridding_true_grid = b200nufft_gridding, the y2k stage of the API (one more pass)
    kern['gridding'] = timed(lambda: lib.b200nufft_gridding_modulated(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st()), kit, 3) / kit
    kern['memset_grid'] = timed(lambda: grid.zero_(), kit, 3) / kit
    kern['ifft_crop'] = timed(lambda: lib.b200nufft_ifft_crop_modulated(A._plan, P(grid.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st()), kit, 3) / kit
    kern['gridding_true_grid'] = timed(lambda: lib.b200nufft_gridding(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st()), kit, 3) / kit
    peak, peak_src = measured_peak()
    dom = 'gridding' if kern['gridding'] >= kern['interp'] else 'interp'
    achieved = ALGO_BYTES / (kern[dom] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dom)
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes': ALGO_BYTES,
                'interp_GBps': ALGO_BYTES / (kern['interp'] * 1e-3) / 1e9,
                'gridding_GBps': ALGO_BYTES / (kern['gridding'] * 1e-3) / 1e9,
                'gridding_true_grid_GBps': ALGO_BYTES / (kern['gridding_true_grid'] * 1e-3) / 1e9,
                'interp_true_grid_GBps': ALGO_BYTES / (kern['interp_true_grid'] * 1e-3) / 1e9}
    del grid, yv, xo

    # ---- extra: configuration 5 (BASELINE configs[4]): 3-D 128^3, 32 coils sharded by coil over the N ranks, k-space CG
    # with the dot products all-reduced; ms per iteration from the difference of two runs (setup excluded) ----
    config5 = None
    plan_bytes = int(lib.b200nufft_plan_bytes(A._plan))
    if not args.no_extras:
        from pynufft_b200.dist import CoilShardedNUFFT, shard_coils
        A.release()
        total_coils = 32
        sl = shard_coils(total_coils, world, rank)
        nloc = sl.stop - sl.start
        A5 = pynufft_b200.NUFFT(dev)
        A5.plan(om, ND, KD, JD, batch=nloc)
        # one rank: the same solver without collectives (CoilShardedNUFFT needs a process group)
        solve5 = (lambda it: A5._solve_device(y5, 'cg', maxiter=it)) if dist is None else \
            (lambda it, op=CoilShardedNUFFT(A5, total_coils): op.solve_cg(y5, maxiter=it))
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        y5 = torch.view_as_complex(torch.randn((M, nloc, 2), generator=g, device=dev))

        def cg_ms(iters):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solve5(iters)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        cg_ms(1)                                                 # warm-up (allocations, cuFFT / attribute set-up)
        lo, hi = 2, 8
        t_lo, t_hi = cg_ms(lo), cg_ms(hi)
        config5 = {'coils_total': total_coils, 'coils_per_gpu': nloc, 'ms_per_iter': (t_hi - t_lo) / (hi - lo),
                   'iters_timed': hi - lo, 'note': 'k-space CG (solve_device.py:351-481), dot products all-reduced over the '
                   'coil shards; device-resident, CUDA events, max over ranks'}
        del y5, solve5
        A5.release()

    # ---- extra: configuration 2 (2-D 256^2 / 512^2 / 6^2, 32 coils, golden-angle radial 402 x 512) stage times ----
    config2 = None
    if not args.no_extras and world == 1:
        th = numpy.arange(402) * numpy.pi * (numpy.sqrt(5.0) - 1.0) / 2.0
        r = numpy.pi * (numpy.arange(512) - 256) / 256
        om2 = numpy.stack([numpy.outer(numpy.cos(th), r), numpy.outer(numpy.sin(th), r)], -1).reshape(-1, 2)
        A2 = pynufft_b200.NUFFT(dev)
        A2.plan(om2, (256, 256), (512, 512), (6, 6), batch=32)
        s2 = torch.view_as_complex(torch.randn((256, 256, 2), device=dev))
        y2 = A2.forward_one2many(s2)
        k2 = A2._y2k_device(y2)
        k2s = A2._grid_storage(k2)[0]
        M2 = om2.shape[0]
        algo2 = 8 * 32 * 512 * 512 + 12 * M2 * 12 + 8 * 32 * M2
        c2t = {}
        c2t['interp'] = timed(lambda: lib.b200nufft_interp(A2._plan, P(k2s.data_ptr()), P(y2.data_ptr()), 32, st()), 30, 3) / 30
        c2t['gridding'] = timed(lambda: lib.b200nufft_gridding(A2._plan, P(y2.data_ptr()), P(k2s.data_ptr()), 32, st()), 30, 3) / 30
        if A2._kspace_modulated(32):    # the same two stages on the phase-modulated grid forward / adjoint / CG keep
            c2t['interp_modulated_grid'] = timed(lambda: lib.b200nufft_interp_modulated(A2._plan, P(k2s.data_ptr()), P(y2.data_ptr()), 32, st()), 30, 3) / 30
            c2t['gridding_modulated_grid'] = timed(lambda: lib.b200nufft_gridding_modulated(A2._plan, P(y2.data_ptr()), P(k2s.data_ptr()), 32, st()), 30, 3) / 30
        c2t['forward_one2many'] = timed(lambda: A2.forward_one2many(s2), 30, 3) / 30
        c2t['adjoint_many2one'] = timed(lambda: A2.adjoint_many2one(y2), 30, 3) / 30
        config2 = {'ms': c2t, 'algorithmic_bytes': algo2,
                   'interp_frac_of_peak': algo2 / (c2t['interp'] * 1e-3) / 1e9 / peak,
                   'gridding_frac_of_peak': algo2 / (c2t['gridding'] * 1e-3) / 1e9 / peak}
        A2.release()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = CpuPath(CPU_SAMPLE_M)
        cpu.pair_seconds()
        secs = [cpu.pair_seconds() for _ in range(3)]
        cpu_baseline = {'value': 1.0 / float(numpy.mean(secs)), 'unit': 'pairs/s', 'cores': 1,
                        'cores_available': os.cpu_count(), 'kind': 'port', 'sample': cpu.describe()}

    if rank == 0:
        line = {
            'metric': 'NUFFT forward+adjoint pairs/s', 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': make_config(world),
            'clocks': clk,
            'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / e2e_iters, 'mode': e2e_mode,
                    'blocking_ms_per_step': (ms_blk / blk_iters) if ms_blk else None,
                    'blocking_value': (blk_iters / (ms_blk * 1e-3)) if ms_blk else None},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'kernel_ms': kern,
            'sustained': sustained,
            'config5_cg': config5,
            'config2': config2,
            'plan_seconds': plan_s, 'plan_bytes': plan_bytes,
            'cpu_baseline': cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def ctypes_ptr(v):
    import ctypes
    return ctypes.c_void_p(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None, help='default: 200 (GPU arm), 10 (reference arm: ~6 s of CPU each)')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extras', action='store_true', help='skip the configuration-5 / configuration-2 extra keys')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        if args.steps is None:
            args.steps = 10
        run_reference(args, rank)
        return
    if args.steps is None:
        args.steps = 200
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
