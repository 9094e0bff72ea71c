#!/usr/bin/env python
"""bench.py -- dense thin-wall inductance (L) build: element-pair integrals per second.

One "step" = one complete build of this rank's row block of L for a synthetic tokamak-vessel
mesh (torus with rectangular ports, jittered vertices; openfusiontoolkit_b200.ThinCurr.meshing).
N=1 builds the whole matrix of a ~20k-vertex vessel (BASELINE.json configs[1]); at N>1 the mesh
grows so that every GPU keeps the N=1 pair count (weak scaling), rows are sharded over the ranks
(symmetric partition: every rank builds the upper trapezoid of its row block, so no pair integral is
evaluated on two devices), no traffic during assembly, and the transposed blocks are exchanged once
afterwards (one NCCL all-to-all, inside the timed step).

  value  : whole-job pair-integrals/s with the model resident in HBM, rows left in HBM
           (pairs = ordered triangle pairs the reference loop nest visits, thin_wall.F90:1028-1035)
  e2e    : same metric through the host-buffer C-ABI call (thincurr_b200_Lmat_shard_host): host mesh
           -> plan upload -> build -> rows copied to pinned host memory, all inside the timed region
  roofline: FP64 pipe.  achieved = algorithmic flops of the reference loop nest (SURVEY.md 8d flop
           model x the measured order histogram) / duration of the tile kernel (device globaltimer, first
           CTA start -> last CTA end, measured live); peak = DFMA micro-benchmark measured in this run
           (MEASURED_PEAKS.json holds no FP64 figure); traffic = DRAM bytes of the kernel from the
           committed ncu --set full capture (profiles/r01_ncu_traffic.json)
  cpu_baseline: the oracle's C/OpenMP restatement of the reference loop (reference flags -O2, same
           schedule) on the host cores, on a bounded row sample of the same mesh

`--impl reference` times that CPU path alone on this arm's config (the Fortran reference cannot be
compiled in this image: no Fortran compiler, no HDF5).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def vessel_dims(n_gpus, base=(100, 200)):
    f = float(n_gpus) ** 0.25
    return int(round(base[0] * f)), int(round(base[1] * f))


def make_mesh(n_gpus, workload):
    from openfusiontoolkit_b200.ThinCurr.meshing import build_torus_vessel
    if workload == 'vessel100k':
        nt, nphi = 224, 448
    elif workload == 'vessel150k':
        nt, nphi = 274, 548
    else:
        nt, nphi = vessel_dims(n_gpus)
    m = build_torus_vessel(nt, nphi, R0=1.0, a=0.5, kappa=1.0, nports=10, jitter=0.05, seed=1234)
    m['dims'] = (nt, nphi)
    return m


def flop_model(hist):
    """algorithmic flops of the reference loop nest for an order histogram."""
    qnp = {4: 6, 5: 7, 6: 12, 7: 15, 8: 16, 9: 19, 10: 25, 11: 28, 12: 33, 13: 46, 14: 46, 15: 55, 16: 55, 17: 72, 18: 72}
    tot = 0.0
    for q in range(4, 19):
        n = qnp[q]
        f = (12 * n * n + 15 * n + 2) if q <= 10 else (247 * n + 1)
        tot += float(hist[q]) * (86 + 63 + f)
    return tot


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonSwPowerCap if hasattr(nv, 'nvmlClocksEventReasonSwPowerCap') else 0x4: 'sw_power_cap',
                 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown'}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['nvml unavailable']}
        return {'sm_mhz': float(np.median(self.sm)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_sample(mesh, seconds_target=15.0, steps=1):
    """Oracle C/OpenMP loop (reference schedule) over a bounded block of row cells; returns
    (pairs/s, threads, description)."""
    from oracle import tw_oracle as tw
    O = tw.OracleModel(mesh['r'], mesh['lc'], None, nodesets=mesh['nodesets'], closures=mesh['closures'])
    L = tw.lib()
    nthreads = int(L.tco_num_threads())
    nrows = min(O.nc, 100 * nthreads)          # one schedule(dynamic,100) chunk per thread
    i0 = (O.nc - nrows) // 2
    out = np.zeros((O.nelems, O.nelems))  # calloc-backed: only the rows the sample touches are ever committed
    best = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        O.compute_Lmat(i0, i0 + nrows, finalize=False, out=out)
        dt = time.perf_counter() - t0
        best = max(best, O.visited / dt)
        last = (O.visited, dt)
    desc = 'rows [%d,%d) of %d row cells x all %d column cells (%d visited pairs in %.1f s), gcc -O2 -fopenmp schedule(dynamic,100)' % (
        i0, i0 + nrows, O.nc, O.nc, last[0], last[1])
    return best, nthreads, desc, last


def run_reference(args, rank, world):
    if rank != 0:
        return
    mesh = make_mesh(args.gpus, args.workload)
    vals = []
    for s in range(args.warmup + args.steps):
        v, nthreads, desc, last = cpu_sample(mesh, steps=1)
        if s >= args.warmup:
            vals.append((v, last[1]))
    value = float(np.mean([v for v, _ in vals]))
    line = {'impl': 'reference', 'metric': 'L-matrix pair-integrals/s', 'value': value, 'unit': 'pairs/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean([t for _, t in vals])),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(mesh, args), 'np': int(mesh['r'].shape[0]), 'nc': int(mesh['lc'].shape[0])},
            'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': nthreads, 'kind': 'port', 'sample': desc},
            'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def ncu_traffic(mesh):
    """dram__bytes_read.sum + dram__bytes_write.sum of lmat_tile_kernel per launch from the committed ncu --set full
    capture of this workload (profiles/r01_ncu_traffic.json), or None when the capture is for another mesh."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')) as f:
            d = json.load(f)
        if int(d['np']) == int(mesh['r'].shape[0]) and int(d['nc']) == int(mesh['lc'].shape[0]):
            return float(d['dram_bytes_read']) + float(d['dram_bytes_write'])
    except Exception:
        pass
    return None


def workload_name(mesh, args):
    return 'synthetic tokamak vessel with 10 ports, %dx%d grid (%d vertices / %d triangles), self-inductance L' % (
        mesh['dims'][0], mesh['dims'][1], mesh['r'].shape[0], mesh['lc'].shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='auto', help='auto (weak-scaled ~20k-vertex vessel) | vessel100k | vessel150k')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU path)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from openfusiontoolkit_b200 import OFT_env
    from openfusiontoolkit_b200 import _interface as I
    from openfusiontoolkit_b200.ThinCurr import ThinCurr

    mesh = make_mesh(args.gpus, args.workload)
    T = ThinCurr(OFT_env(nthreads=-1))
    T.setup_model(r=mesh['r'], lc=mesh['lc'], nodesets=mesh['nodesets'], closures=mesh['closures'])
    N = T.nelems
    # N > 1: symmetric partition -- every rank builds the upper trapezoid of its row block (no pair integral is
    # evaluated on two devices) and the transposed blocks are exchanged once after the assembly (NCCL send/recv)
    sym = world > 1
    ids_all = [T.shard_rows_sym(world, s) for s in range(world)] if sym else None
    rows = ids_all[rank] if sym else T.shard_rows(world, rank)
    nrows = len(rows)
    out = torch.empty((nrows, N), dtype=torch.float64, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # reference-loop statistics of the whole mesh (outside the timed region)
    hist, visited = T.pair_stats()
    flops_total = flop_model(hist)
    # this rank's share of the algorithmic work: rows are balanced by cell count
    share = 1.0 / world

    def step(stats=False):
        if sym:
            st_ = T.compute_Lmat_shard_sym(world, rank, out, stream=stream, stats=stats)
            T.exchange_symmetric(out, world, rank, row_ids=ids_all)
            return st_
        return T.compute_Lmat_shard(world, rank, out, stream=stream, stats=stats)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = I.b200_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t0 = time.perf_counter()
    ev[0].record()
    for s in range(args.steps):
        step()
        ev[s + 1].record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.result()
    launches = I.b200_launch_count() - launches0
    ms_dev = ev[0].elapsed_time(ev[args.steps])
    tt = torch.tensor([ms_dev, wall * 1e3], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_dev, ms_wall = [float(v) for v in tt.tolist()]
    ms_step = ms_dev / args.steps
    value = visited / (ms_step * 1e-3)

    # kernel-only time of one step (events right around the tile kernel would need hooks inside the
    # library; the step is memset + one tile kernel, so time a step without the memset share)
    st = step(stats=True)
    # lmat_tile_kernel's own duration from the device's globaltimer (first CTA start -> last CTA end, written by the
    # kernel when stats are requested); the step additionally holds the output memset, the row-map kernel and the
    # symmetrisation pass
    kern_ms = (int(st[4]) - int(st[6])) * 1e-6
    if not (0.0 < kern_ms <= 1.05 * ms_step):
        kern_ms = ms_step

    # N > 1: sanity check of the exchange outside the timed region -- block sums of L[rows r][DOFs s] and
    # L[rows s][DOFs r] (its transpose, held by the other rank) must agree
    exchange_check = None
    if sym:
        bs = torch.stack([out[:, torch.as_tensor(i.astype(np.int64), device='cuda')].sum() for i in ids_all])
        allbs = [torch.empty_like(bs) for _ in range(world)]
        dist.all_gather(allbs, bs)
        M = torch.stack(allbs).cpu().numpy()
        exchange_check = float(np.abs(M - M.T).max() / np.abs(M).max())

    # e2e through the host-buffer entry point
    e2e = None
    if not args.no_e2e:
        host = torch.empty((len(T.shard_rows(world, rank)), N), dtype=torch.float64, pin_memory=True)
        hnp = host.numpy()
        est = None
        for _ in range(2):
            est = T.compute_Lmat_shard_host(world, rank, hnp, stats=True)
        barrier()
        t0 = time.perf_counter()
        nrep = max(1, min(args.steps, 3))
        for _ in range(nrep):
            est = T.compute_Lmat_shard_host(world, rank, hnp, stats=True)
        barrier()
        dt = (time.perf_counter() - t0) / nrep
        tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {'value': visited / dt, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(est[5]) * world, 'd2h_bytes_per_step': int(est[6]) * world,
               'ms_per_step': dt * 1e3, 'api': 'thincurr_b200_Lmat_shard_host (host mesh -> pinned host rows)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_clock = np.zeros(1)
    import ctypes
    peak_tf = float(I.b200_dfma_peak(local_rank, peak_clock.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    achieved_tf = flops_total * share / (kern_ms * 1e-3) / 1e12
    roofline = {'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf if peak_tf > 0 else None,
                'traffic': ncu_traffic(mesh), 'peak_source': 'DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)',
                'algorithmic_flops_per_step': flops_total * share, 'flops_per_pair': flops_total / visited,
                'hbm_write_GBps': nrows * N * 8 / (kern_ms * 1e-3) / 1e9, 'kernel': 'lmat_tile_kernel',
                'kernel_ms': kern_ms, 'cta_finish_spread': {'first_ms': (int(st[7]) - int(st[6])) * 1e-6, 'last_ms': (int(st[4]) - int(st[6])) * 1e-6},
                'device_evals': {'far_pairs': int(st[0]), 'near_T': int(st[1]), 'inv_r': int(st[2]), 'phipot': int(st[3])}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, nthreads, desc, _ = cpu_sample(mesh)
        cpu = {'value': v, 'unit': 'pairs/s', 'cores': nthreads, 'kind': 'port', 'sample': desc}
    line = {'metric': 'L-matrix pair-integrals/s', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(mesh, args), 'np': int(mesh['r'].shape[0]), 'nc': int(mesh['lc'].shape[0]),
                       'nelems': int(N), 'visited_pairs': int(visited), 'nc2_pairs': int(mesh['lc'].shape[0]) ** 2,
                       'order_hist': {str(q): int(hist[q]) for q in range(4, 19)}, 'sharding': ('row blocks, 1 shard, no collective' if not sym else
                                    'row blocks balanced over the upper trapezoid, %d shards, no traffic during assembly; transposed blocks '
                                    'exchanged once afterwards (NCCL send/recv, inside the timed step)' % world),
                       'exchange_check_rel': exchange_check,
                       'l2': 'flushed every step by the %.1f GB output memset' % (nrows * N * 8 / 1e9), 'plan': T.plan_info()},
            'wall_ms_per_step': ms_wall / args.steps, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
