#!/usr/bin/env python
"""
Benchmark of the hot path: model evaluations per second for the 4-TLS (d = 16) library
(BASELINE.json `metric`, workload `configs[2]`: 1 qubit + 3 defects, 4096 chains per GPU, 200 evaluation
times, 3 observables).  One "step" = one pass of the hot path over one batch: every chain's current proposal
(4096 parameter vectors per GPU) -> Lindblad dynamics at the target times -> observables -> loss.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA kernels)
    python bench.py --impl reference [...]                         # the reference's CPU path (emulated qutip mesolve)

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how every field is measured.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OBS3 = ['sigmax', 'sigmay', 'sigmaz']
N_TIMES = 200
T_SPAN = (0.0, 2.5)
CONFIG_INDEX = 11          # configs.py "EVERYTHING" library


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--chains', type=int, default=4096, help='chains (models per step) per GPU')
    ap.add_argument('--defects', type=int, default=3)
    ap.add_argument('--qubits', type=int, default=1)
    ap.add_argument('--kernel', default='auto', choices=['auto', 'generic', 'warp_mma', 'warp_cheb'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-chains', action='store_true', help='skip the RJMCMC steps/s measurement')
    ap.add_argument('--cpu-sample', type=int, default=0, help='models in the CPU-baseline sample (0 = 64 per core, in both CPU legs)')
    ap.add_argument('--workload', default='eval', choices=['eval', 'multiset'],
                    help="eval: the headline evaluation batch (BASELINE.json configs[2]); multiset: configs[4], a fixed sweep of "
                         "8192 chains over mixed d = 4/8/16 models on N GPUs (strong scaling), all_gather at checkpoints")
    ap.add_argument('--sweep-chains', type=int, default=8192, help='multiset: chains of the whole sweep (fixed as N grows)')
    ap.add_argument('--sweep-weak', action='store_true', help='multiset: --sweep-chains chains PER GPU (weak scaling) instead of in total')
    ap.add_argument('--checkpoint-every', type=int, default=50, help='multiset: steps between checkpoint all_gathers')
    ap.add_argument('--sustain-seconds', type=float, default=2.0, help='length of the sustained-throughput measurement (0 = skip)')
    ap.add_argument('--chain-steps', type=int, default=1000, help='RJMCMC steps measured after --chain-burn-in (independent of --steps)')
    ap.add_argument('--chain-burn-in', type=int, default=200)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# shared workload (product-side generator; tests check it equals the oracle's random_library_spec)
def make_layout(args):
    from effective_model_learning_b200 import configs, synthetic
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[CONFIG_INDEX])
    return synthetic.LibraryLayout(args.qubits, args.defects, hyper), hyper


def workload_name(args):
    d = 2 ** (args.qubits + args.defects)
    return (f'{args.qubits} qubit + {args.defects} defects (d={d}), {args.chains} chains/GPU, '
            f'{N_TIMES} times on [0,2.5], observables sx,sy,sz, library config {CONFIG_INDEX}')


# ----------------------------------------------------------------------------------------------------
# CPU reference path: the reference evaluates one model per qutip.mesolve call, chains as parallel OS
# processes (launch_execute_learning_multiset.sh:60-63).  qutip is not installable, so the timed code is
# the oracle's restatement of qutip 4.7.5 mesolve (CSR Liouvillian + ZVODE adams/12, atol 1e-8, rtol 1e-6).
def probe_real_reference():
    """
    Tier A of BASELINE.md: the UNMODIFIED reference (basic_model.BasicModel.calculate_dynamics -> qutip.mesolve,
    basic_model.py:381-388) is timed when both qutip 4.x and the reference tree are present.  Returns
    (usable, report) -- the report goes into the JSON line either way, so that every line states which CPU arm ran.
    """
    report = {}
    ref_path = os.environ.get('EML_REFERENCE_PATH', '/root/reference')
    report['reference_tree'] = ref_path if os.path.isfile(os.path.join(ref_path, 'basic_model.py')) else \
        f'absent ({ref_path}/basic_model.py does not exist on this box)'
    try:
        import qutip
        if not hasattr(qutip, 'mesolve') or 'qutip_shim' in (getattr(qutip, '__file__', '') or ''):
            raise ImportError('only the test shim is importable')
        report['qutip_import'] = f'ok, version {getattr(qutip, "__version__", "?")}'
        ok = True
    except Exception as err:      # noqa: BLE001
        report['qutip_import'] = f'{type(err).__name__}: {err}'
        ok = False
    return ok and not report['reference_tree'].startswith('absent'), report


def _cpu_worker(rows):
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import lindblad_oracle as lo
    from effective_model_learning_b200 import configs, synthetic
    (Q, D, rows, targets, real_reference) = rows
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[CONFIG_INDEX])
    layout = synthetic.LibraryLayout(Q, D, hyper)
    ts = np.linspace(*T_SPAN, N_TIMES)
    out = []
    if real_reference:
        # the reference's own classes and its own calculate_dynamics (one qutip.mesolve per model), unmodified
        sys.path.insert(0, os.environ.get('EML_REFERENCE_PATH', '/root/reference'))
        import learning_model as ref_learning_model
        for row in rows:
            spec = lo.spec_from_model(layout.model_from_params(row))
            m = ref_learning_model.LearningModel()
            made = [m.add_TLS(is_qubit=t['is_qubit'], energy=t['energy'], couplings={}, Ls=dict(t['Ls']),
                              initial_state=False) for t in spec['tls']]
            for i, t in enumerate(spec['tls']):
                for (j, strength, op_pairs) in t['couplings']:
                    made[i].couplings.setdefault(made[j], []).append((strength, list(op_pairs)))
            import definitions as ref_definitions
            for k, tls in enumerate(m.TLSs):
                tls.initial_state = ref_definitions.ops['plus'] if tls.is_qubit else ref_definitions.ops['mm']
            m.build_operators()
            ex = m.calculate_dynamics(ts, OBS3)
            out.append(lo.total_dev(ex, targets, N_TIMES))
        return out
    for row in rows:
        spec = lo.spec_from_model(layout.model_from_params(row))
        ex = lo.expect_qutip_emulation(spec, ts, OBS3)
        out.append(lo.total_dev(ex, targets, N_TIMES))
    return out


class CpuPool:
    def __init__(self):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        for v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
            os.environ[v] = '1'
        self.pool = mp.get_context('spawn').Pool(self.cores)

    def evaluate(self, Q, D, rows, targets, real_reference=False):
        chunks = [c for c in np.array_split(np.asarray(rows), self.cores) if len(c)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(Q, D, c, targets, real_reference) for c in chunks])
        dt = time.perf_counter() - t0
        return dt, [x for r in res for x in r]

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_targets(args, layout):
    """Targets for the CPU arm: exact observables of the ground-truth model + N(0, 0.01^2) (oracle allowed here)."""
    from oracle import lindblad_oracle as lo
    truth = layout.random_params(np.random.default_rng(20260000 + 10 * args.qubits + args.defects))
    ts = np.linspace(*T_SPAN, N_TIMES)
    ex = lo.expect_exact(lo.spec_from_model(layout.model_from_params(truth)), ts, OBS3)
    noise = np.random.default_rng(4242)
    return [np.asarray(e) + noise.normal(0, 0.01, N_TIMES) for e in ex]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    layout, _ = make_layout(args)
    targets = cpu_targets(args, layout)
    real, probe = probe_real_reference()
    pool = CpuPool()
    # 64 models per core and step (~1.2 s of wall clock per step): the same sample as the in-line cpu_baseline of the GPU
    # arm, so that the two CPU figures agree and the pool dispatch does not flatter the GPU/CPU ratio
    per_step = args.cpu_sample or 64 * pool.cores
    per_step = min(per_step, args.chains)
    rows = layout.batch_params(1000 + np.arange(per_step))
    for _ in range(max(1, min(args.warmup, 1))):
        pool.evaluate(args.qubits, args.defects, rows[:pool.cores], targets, real)
    t_total = 0.0
    for _ in range(args.steps):
        dt, _ = pool.evaluate(args.qubits, args.defects, rows, targets, real)
        t_total += dt
    pool.close()
    value = per_step * args.steps / t_total
    sample = f'{per_step} of the {args.chains} models of the batch per step (seeds 1000..{1000 + per_step - 1})'
    line = {
        'impl': 'reference', 'metric': 'model evaluations/sec (4-TLS, d=16)' if (args.qubits, args.defects) == (1, 3)
        else 'model evaluations/sec', 'value': value, 'unit': 'evals/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'algorithm': (
            'the unmodified reference: BasicModel.calculate_dynamics -> qutip.mesolve, one OS process per core' if real else
            'emulated qutip 4.7.5 mesolve: CSR Liouvillian + scipy ZVODE adams/12 atol 1e-8 rtol 1e-6, one OS process per '
            'core (the real reference could not run here: see reference_probe)')},
        'cpu_baseline': {'value': value, 'unit': 'evals/s', 'cores': pool.cores, 'kind': 'reference' if real else 'port',
                         'sample': sample},
        'reference_probe': probe,
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'rjmcmc_equivalent_steps_per_s': value / 3.0,   # the reference spends 3 evaluations per RJMCMC step
    }
    emit_line(line)


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """
    SM clock, power and clock-event reasons sampled DURING the timed region: NVML in a background thread every
    20 ms (pynvml), falling back to an `nvidia-smi -lms 100` subprocess when NVML cannot be loaded.
    """
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, device):
        self.rows = []          # (t, sm_mhz, sm_max_mhz, power_w, [reasons])
        self.proc = None
        self.stop_flag = False
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            index = int(visible.split(',')[device]) if visible and visible.split(',')[device].isdigit() else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = 'nvml'
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(device)], stdout=subprocess.PIPE, text=True)
            self.source = 'nvidia-smi'
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = ((n.nvmlClocksEventReasonHwSlowdown, 'hw_slowdown'), (n.nvmlClocksEventReasonHwThermalSlowdown, 'hw_thermal_slowdown'),
                (n.nvmlClocksEventReasonSwThermalSlowdown, 'sw_thermal_slowdown'), (n.nvmlClocksEventReasonSwPowerCap, 'sw_power_cap'))
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append((time.perf_counter(), sm, self.smax, pw, [name for bit, name in bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.02)

    def _read_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(',')]
            try:
                self.rows.append((time.perf_counter(), float(parts[0]), float(parts[1]), float(parts[2]),
                                  [n for n, p in zip(self.NAMES, parts[3:7]) if p.lower().startswith('active')]))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        if self.source is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no NVML / nvidia-smi available']}
        time.sleep(0.05 if self.source == 'nvml' else 0.15)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        inside = [r for r in self.rows if t0 - 0.02 <= r[0] <= t1 + 0.02] or self.rows[-3:]
        sm = [r[1] for r in inside]
        reasons = sorted({name for r in inside for name in r[4]})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max((r[2] for r in inside), default=None),
                'power_w_max': max((r[3] for r in inside), default=None), 'samples': len(sm), 'reasons': reasons,
                'source': self.source}


def canonical_dense_flops(n_tls, T, K):
    """Pade-13 scaling-and-squaring expm (s=0) + T-1 mat-vecs + traces, N = d^2."""
    N = 4 ** n_tls
    return 8 * N ** 3 * 6 + (32 / 3) * N ** 3 + 8 * N * N * (T - 1) + 8 * N * K * T + 3 * K * T


def run_b200(args):
    import torch
    import torch.distributed as dist
    from effective_model_learning_b200 import _native, engine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    kernel = {'auto': _native.KERNEL_AUTO, 'generic': _native.KERNEL_GENERIC, 'warp_mma': _native.KERNEL_WARP_MMA,
              'warp_cheb': _native.KERNEL_WARP_CHEB}[args.kernel]

    layout, _ = make_layout(args)
    ts = np.linspace(*T_SPAN, N_TIMES)
    ctx = layout.context(ts, OBS3, device=local, kernel=kernel)
    # targets: ground-truth model evaluated by the kernels themselves (parity with the oracle is what the
    # tests establish) + N(0, 0.01^2), execute_learning_full.py:114
    truth = layout.random_params(np.random.default_rng(20260000 + 10 * args.qubits + args.defects))
    ex, _ = ctx.evaluate(truth[None], want_expect=True)
    noise = np.random.default_rng(4242)
    targets = ex[0].real + noise.normal(0, 0.01, (3, N_TIMES))
    ctx.set_targets(targets)
    plan = ctx.plan

    B = args.chains
    params_host = layout.batch_params(1000 + rank * B + np.arange(B))         # this rank's chains
    params = torch.from_numpy(params_host).to(dev)
    loss = torch.zeros(B, dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    # Inputs larger than L2: every step evaluates a DIFFERENT parameter matrix out of a ring whose total size is 1.3x the
    # L2 capacity the device reports (B200: 126 MB), as every MCMC step brings new proposals -- no step finds its input
    # in L2 and no flush kernel sits in the timed region.  The ring entries are row permutations of the batch (same
    # models, same cost distribution, different addresses and a different order every step).
    l2_bytes = int(torch.cuda.get_device_properties(dev).L2_cache_size)
    ring_n = max(2, int(np.ceil(1.3 * l2_bytes / (B * layout.n_params * 8))))
    gen = torch.Generator(device='cpu').manual_seed(7 + rank)
    ring = torch.empty((ring_n, B, layout.n_params), dtype=torch.float64, device=dev)
    ring[0] = params
    for r in range(1, ring_n):
        ring[r] = params[torch.randperm(B, generator=gen).to(dev)]
    ring_bytes = ring.numel() * 8
    # the flush used by round 1 is kept for the comparison figure roofline.kernel_ms_after_l2_flush
    flush_bytes = max(int(1.25 * l2_bytes), 64 * 1024 * 1024)
    flush = torch.empty(flush_bytes // 4, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ring_pos = [0]

    def step():
        ring_pos[0] = (ring_pos[0] + 1) % ring_n
        plan.eval_device(B, ring[ring_pos[0]].data_ptr(), 0, 0, loss.data_ptr(), status.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # the checkpoint gather of the timed region, once untimed as well (NCCL sets its channels up on the first call of a kind)
    gathered = torch.empty((world, B), dtype=torch.float64, device=dev) if world > 1 else None
    if world > 1:
        dist.all_gather_into_tensor(gathered, loss)
    barrier()
    st = status.cpu().numpy()
    if (st < 0).any():
        raise RuntimeError('a model was reported as not converged during warm-up')
    applications = float(st.sum())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.1)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        k_ev[i][0].record()
        step()
        k_ev[i][1].record()
    if world > 1:
        # the path's only collective: per-chain statistics gathered at a checkpoint (once per timed region)
        dist.all_gather_into_tensor(gathered, loss)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _native.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    elapsed_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    if world > 1:
        t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kernel_ms = float(t[0]), float(t[1])
    value = B * world * args.steps / (elapsed_ms * 1e-3)
    # comparison figure: the round-1 protocol (one fixed input, L2 flushed by a 1.25 x L2 memset before every step)
    f_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 20))]
    for a_, b_ in f_ev:
        flush.zero_()
        a_.record()
        plan.eval_device(B, params.data_ptr(), 0, 0, loss.data_ptr(), status.data_ptr(), stream)
        b_.record()
    torch.cuda.synchronize()
    kernel_ms_flushed = float(np.mean([a_.elapsed_time(b_) for a_, b_ in f_ev]))

    # ---- end-to-end through the public host API: params in page-locked host memory in, host loss out, every step
    params_pinned = torch.from_numpy(params_host).pin_memory().numpy()
    loss_pinned = torch.empty(B, dtype=torch.float64).pin_memory().numpy()
    for _ in range(2):
        ctx.evaluate(params_pinned, want_expect=False, want_loss=True, loss_out=loss_pinned)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, loss_host = ctx.evaluate(params_pinned, want_expect=False, want_loss=True, loss_out=loss_pinned)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = B * world * args.steps / e2e_s
    if not np.allclose(loss_host, loss.cpu().numpy(), rtol=0, atol=0):
        raise RuntimeError('host-API and device-API losses differ')

    # ---- roofline of the (single) kernel: FP64 FLOPs it executes / CUDA-event time, vs measured FP64 peak
    fp64_fma = _native.measure_fp64_peak(local, False, 0.3)
    fp64_mma = _native.measure_fp64_peak(local, True, 0.3)
    fp64_mixed = _native.measure_fp64_peak(local, 2, 0.3)      # DMMA and DFMA interleaved in every warp
    peak = max(fp64_fma, fp64_mma)
    flops_per_app = engine.flops_per_application(plan.kernel_name, layout.n_tls)
    flops_launch = flops_per_app * applications
    achieved = flops_launch / (kernel_ms * 1e-3)
    canon = canonical_dense_flops(layout.n_tls, N_TIMES, 3)
    traffic = None
    traffic_note = None
    ncu_pipes = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            entry = json.load(f).get(f'{plan.kernel_name}:n_tls={layout.n_tls}:chains={B}')
        if entry:
            traffic = entry['bytes']
            traffic_note = entry.get('note')
            ncu_pipes = {k: entry[k] for k in ('fp64_pipe_pct', 'tensor_fp64_pct', 'issue_pct', 'duration_ms') if k in entry}
    except (OSError, ValueError):
        pass
    roofline = {
        'bound': 'tensor', 'pipe': 'fp64 (DFMA/DMMA)', 'kernel': plan.kernel_name,
        'achieved': achieved / 1e12, 'peak': peak / 1e12, 'unit': 'TFLOP/s', 'frac': achieved / peak,
        'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu --set full capture of the same workload)',
        'traffic_note': traffic_note,
        'peak_source': 'measured in this run: eml_measure_fp64_peak (register-resident DFMA / DMMA m8n8k4 chains); '
                       'MEASURED_PEAKS.json has no FP64 figure',
        'peak_dfma_tflops': fp64_fma / 1e12, 'peak_dmma_tflops': fp64_mma / 1e12,
        'peak_mixed_dmma_dfma_tflops': fp64_mixed / 1e12,
        'flops_per_launch': flops_launch, 'generator_applications_per_eval': applications / B,
        'flops_per_application': flops_per_app, 'kernel_ms': kernel_ms, 'kernel_ms_after_l2_flush': kernel_ms_flushed,
        'canonical_dense_gflop_per_eval': canon / 1e9,
        'canonical_dense_equivalent_tflops': canon * B / (kernel_ms * 1e-3) / 1e12,
        'algorithmic_hbm_bytes_per_eval': 8 * layout.n_params + 8 + 4,
        # pipe utilisation of the same kernel and workload from the committed ncu capture (percent of peak over the launch;
        # the FP64 pipe and the FP64 tensor path share one datapath) -- not measured in this run
        'ncu': ncu_pipes,
    }

    # ---- sustained throughput: >= sustain_seconds of back-to-back launches with the clocks sampled INSIDE the region
    #      (the timed region above is a burst of a few milliseconds).  Two regimes: the headline batch repeated without
    #      an L2 flush, and 16 x the batch per launch (65536 chains: the regime long chain runs live in).
    if args.sustain_seconds > 0:
        sustained = {}
        peak_sustained = _native.measure_fp64_peak(local, True, args.sustain_seconds)
        reps_big = 16
        params_big = params.repeat(reps_big, 1).contiguous()
        loss_big = torch.zeros(B * reps_big, dtype=torch.float64, device=dev)
        status_big = torch.zeros(B * reps_big, dtype=torch.int32, device=dev)
        for label, (n_models, p_, l_, s_) in (('headline_batch_back_to_back', (B, params, loss, status)),
                                              (f'{reps_big}x_batch_per_launch', (B * reps_big, params_big, loss_big, status_big))):
            for _ in range(2):
                plan.eval_device(n_models, p_.data_ptr(), 0, 0, l_.data_ptr(), s_.data_ptr(), stream)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            plan.eval_device(n_models, p_.data_ptr(), 0, 0, l_.data_ptr(), s_.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            n_launch = max(3, int(np.ceil(args.sustain_seconds * 1e3 / a.elapsed_time(b))))
            smp = ClockSampler(local) if rank == 0 else None
            if smp:
                time.sleep(0.1)
            barrier()
            w0 = time.perf_counter()
            a.record()
            for _ in range(n_launch):
                plan.eval_device(n_models, p_.data_ptr(), 0, 0, l_.data_ptr(), s_.data_ptr(), stream)
            b.record()
            barrier()
            w1 = time.perf_counter()
            ms = a.elapsed_time(b)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            apps = float(s_.sum().item())
            sustained[label] = {'models_per_launch': n_models, 'launches': n_launch, 'seconds': ms * 1e-3,
                                'evals_per_s': n_models * world * n_launch / (ms * 1e-3),
                                'frac_of_sustained_dmma_peak': flops_per_app * apps * n_launch / (ms * 1e-3) / peak_sustained,
                                'clocks': smp.stop(w0, w1) if smp else None}
        sustained['peak_dmma_tflops_sustained'] = peak_sustained / 1e12
        sustained['note'] = (f'>= {args.sustain_seconds:g} s of back-to-back launches per regime, no L2 flush, parameters resident in HBM; '
                             'the DMMA peak is measured over the same length of time')
        roofline['sustained'] = sustained
        del params_big, loss_big, status_big

    # ---- transparency: the same batch with the steady-sector reduction switched off (the whole density matrix evolves)
    if 'sector' in plan.kernel_name:
        ctx_full = layout.context(ts, OBS3, device=local, kernel=kernel, flags=_native.FLAG_NO_SECTOR_REDUCTION)
        ctx_full.set_targets(targets)
        plan_full = ctx_full.plan
        loss_full = torch.zeros(B, dtype=torch.float64, device=dev)
        for _ in range(2):
            plan_full.eval_device(B, params.data_ptr(), 0, 0, loss_full.data_ptr(), status.data_ptr(), stream)
        n_full = min(args.steps, 20)
        f_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_full)]
        for a, b in f_ev:
            flush.zero_()
            a.record()
            plan_full.eval_device(B, params.data_ptr(), 0, 0, loss_full.data_ptr(), status.data_ptr(), stream)
            b.record()
        torch.cuda.synchronize()
        ms_full = float(np.mean([a.elapsed_time(b) for a, b in f_ev]))
        roofline['without_sector_reduction'] = {
            'kernel': plan_full.kernel_name, 'kernel_ms': ms_full, 'evals_per_s_kernel_only': B / (ms_full * 1e-3),
            'max_abs_loss_diff': float((loss_full - loss).abs().max()),
            'note': 'rho0 of this workload (qubit |+><+|, defects maximally mixed: the reference\'s production state) has a steady '
                    'parity-even half under the Pauli library, which the default kernel does not evolve (DESIGN.md section 4.1)'}

    line = {
        'metric': 'model evaluations/sec (4-TLS, d=16)' if (args.qubits, args.defects) == (1, 3) else 'model evaluations/sec',
        'value': value, 'unit': 'evals/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'kernel': plan.kernel_name,
                   'algorithm': ('Chebyshev (Bessel-coefficient) series of the propagator acting on the d x d density matrix, one '
                                 'recurrence per macro step, Bessel-weighted dense output (DESIGN.md section 4)'
                                 if 'cheb' in plan.kernel_name else
                                 'Taylor action on the d x d density matrix with dense output (DESIGN.md section 4)'),
                   'l2': f'inputs larger than L2: every step reads a different parameter matrix out of a ring of {ring_n} row-permuted '
                         f'copies of the batch ({ring_bytes / 1e6:.0f} MB = 1.3 x the {l2_bytes / 1e6:.0f} MB L2), no flush kernel in the timed '
                         'region; roofline.kernel_ms_after_l2_flush is the round-1 protocol (fixed input, 1.25 x L2 memset before every step)',
                   'partitioning': f'chains sharded contiguously over {world} GPU(s), no data-path collective; '
                                   'one all_gather of per-chain loss per timed region when n_gpus > 1'},
        'e2e': {'value': e2e_value, 'unit': 'evals/s', 'h2d_bytes_per_step': int(B * layout.n_params * 8),
                'd2h_bytes_per_step': int(B * 8 + B * 4),
                'transfer': 'EvalContext.evaluate -> eml_eval_batch_host with page-locked host arrays, every step: the prepare kernel '
                            'reads the parameter matrix from the page-locked host memory itself and the evaluator stores loss and '
                            'status there (unified addressing; the same bytes cross PCIe as with cudaMemcpyAsync, but inside the '
                            'kernels instead of before and after them), then cudaStreamSynchronize; EML_HOST_ZEROCOPY=0 restores '
                            'the explicit copies'},
        'gpu_launches': int(launches),
        'gpu_launches_counted': 'evaluator kernels (eml_launch_count); each Chebyshev evaluator launch follows one prepare_models_kernel '
                                'launch of the same library, so a step is two kernels (launch list under profiles/)',
        'roofline': roofline,
        'clocks': clocks,
    }

    # ---- second headline metric: RJMCMC steps/s of the batched chain engine (propose -> evaluate -> accept on device)
    if not args.no_chains:
        from effective_model_learning_b200 import BatchedChains, configs, definitions
        cfg = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[CONFIG_INDEX])
        rj = {}
        for label, init in (('from_initial_model', None), ('from_library_models', params_host)):
            eng = BatchedChains(ts, targets, OBS3, n_chains=B, n_defects=args.defects, n_qubits=args.qubits,
                                seed=1000 + rank, device=local, initial_params=init,
                                defect_initial_state=definitions.ops['mm'], **cfg)
            # steady state: >= chain_burn_in proposals per chain first, then chain_steps timed (independent of --steps)
            n_burn, n_steps = max(args.chain_burn_in, 2), max(args.chain_steps, 1)
            eng.run(n_burn, stream=stream)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            eng.run(n_steps, stream=stream)
            c1.record()
            barrier()
            ms = c0.elapsed_time(c1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            table = eng.checkpoint(B * world) if world > 1 else eng.read()['stats']
            _, n_bad = eng.bad_evaluations()
            rj[label] = {'steps_per_s': B * world * n_steps / (ms * 1e-3), 'ms_per_step': ms / n_steps,
                         'burn_in_steps': n_burn, 'timed_steps': n_steps,
                         'mean_processes_per_model': float(np.mean(table[:, 7])),
                         'acceptance_rate': float((table[:, 2].sum() + table[:, 4].sum()) / max(1.0, table[:, 3].sum() + table[:, 5].sum())),
                         'median_best_loss': float(np.median(table[:, 1])), 'unconverged_proposals_rank0': n_bad,
                         # what the reference's loop would need for the same steps: 3 evaluations per step (learning_chain.py:559, 782-783)
                         'reference_equivalent_evaluations_per_s': 3.0 * B * world * n_steps / (ms * 1e-3)}
            eng.close()
        rj['note'] = ('one step = one proposal per chain: device-side move proposal, evaluation of the proposal, MH accept/reject; '
                      'from_initial_model starts every chain at make_initial_model(1, D) (SURVEY 8d), from_library_models at the '
                      'random library models of the evaluation batch (same model complexity as the headline metric) -- the '
                      'representative figure; both are measured over chain_steps proposals per chain after chain_burn_in; '
                      'the reference spends 3 evaluations per step (SURVEY F6), i.e. its steps/s = its evaluations/s / 3')
        line['rjmcmc'] = rj

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pool = CpuPool()
        n = args.cpu_sample or 64 * pool.cores         # ~20 core-seconds of CPU work
        n = min(n, B)
        rows = params_host[:n]
        tg = [targets[m] for m in range(3)]
        pool.evaluate(args.qubits, args.defects, rows[:pool.cores], tg)      # warm the workers (imports)
        dt, cpu_loss = pool.evaluate(args.qubits, args.defects, rows, tg)
        pool.close()
        line['cpu_baseline'] = {'value': n / dt, 'unit': 'evals/s', 'cores': pool.cores, 'kind': 'port',
                                'sample': f'first {n} models of the batch, emulated qutip 4.7.5 mesolve '
                                          '(ZVODE adams/12, atol 1e-8, rtol 1e-6), one process per core',
                                'max_abs_loss_diff_vs_gpu': float(np.abs(np.array(cpu_loss) - loss_host[:n]).max())}
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: the multiset sweep (reference launcher launch_execute_learning_multiset.sh:35-63) as one job
def run_multiset(args):
    """
    A FIXED sweep (strong scaling): --sweep-chains chains = 2 synthetic target datasets x library config 11 x D in {1, 2, 3}
    (d = 4, 8, 16) x repetitions, every (dataset, D) group split over the N ranks (cost-balanced: dist.balanced_owners),
    each rank advancing its engines concurrently on separate streams.  One step = one proposal for EVERY chain of the
    sweep; every --checkpoint-every steps the per-chain statistics of all ranks are gathered with one all_gather, inside
    the timed region.  value = sweep chains x steps / time (max over ranks).
    """
    import torch
    import torch.distributed as dist
    from effective_model_learning_b200 import _native, configs, definitions, synthetic
    from effective_model_learning_b200 import dist as edist
    from effective_model_learning_b200.multiset import MultisetSweep

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hyper = configs.get_hyperparams(list(configs.specific_experiment_chain_hyperparams)[CONFIG_INDEX])
    ts = np.linspace(*T_SPAN, N_TIMES)
    datasets = []
    for D in (2, 3):      # two synthetic target datasets, generated from ground-truth models by the kernels themselves
        lay = synthetic.LibraryLayout(1, D, hyper)
        ctx = lay.context(ts, OBS3, device=local)
        ex, _ = ctx.evaluate(lay.random_params(np.random.default_rng(20260000 + 10 + D))[None], want_expect=True)
        datasets.append((ts, ex[0].real + np.random.default_rng(D).normal(0, 0.01, (3, N_TIMES)), OBS3))
    defects = [1, 2, 3]
    groups = len(datasets) * len(defects)
    total = args.sweep_chains * (world if args.sweep_weak else 1)
    reps = [total // groups + (1 if i < (total % groups + 1) // 2 else 0) for i in range(len(defects))]
    sweep = MultisetSweep(datasets, config_numbers=[CONFIG_INDEX], defects_numbers=defects, repetitions_numbers=reps, seed=1,
                          qubit_initial_state=definitions.ops['plus'], defect_initial_state=definitions.ops['mm'], device=local)

    def barrier():
        sweep.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sweep.run(max(args.warmup, 3))
    sweep.checkpoint()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    barrier()
    w0 = time.perf_counter()
    e0.record()
    done, n_ckpt, table = 0, 0, None
    while done < args.steps:
        n = min(args.checkpoint_every, args.steps - done)
        sweep.run(n)
        done += n
        table = sweep.checkpoint()          # synchronises this rank's streams, one all_gather, table on every rank
        n_ckpt += 1
    e1.record()
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1) if sampler else None
    seconds = w1 - w0                      # the engines run on their own streams: wall clock between two full barriers
    if world > 1:
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t[0])
    value = sweep.n_chains * args.steps / seconds
    by_d = {}
    for spec, row in zip(sweep.chains, table):
        by_d.setdefault(spec.defects, []).append(row)
    owned = [len(o) for o in sweep.owners]
    line = {
        'metric': 'RJMCMC steps/sec (multiset sweep, mixed 2-4 TLS)', 'value': value, 'unit': 'chain-steps/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * seconds / args.steps, 'higher_is_better': True,
        'scaling': 'weak' if args.sweep_weak else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'multiset sweep: {sweep.n_chains} chains = {len(datasets)} target datasets x library config {CONFIG_INDEX} x '
                               f'D in {defects} (d = 4, 8, 16) x {reps} repetitions, {N_TIMES} times on [0,2.5], observables sx,sy,sz; '
                               f'checkpoint all_gather every {args.checkpoint_every} steps',
                   'partitioning': f'every (dataset, D) group split over the {world} rank(s) (cost-balanced), chains per rank {owned}; '
                                   'no data-path collective, one all_gather of the [chains][8] statistics table per checkpoint',
                   'l2': 'working set (chain states, < 10 MB per rank) is L2 resident by design; nothing is flushed between steps'},
        'e2e': {'value': value, 'unit': 'chain-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': int(len(sweep.owned) * 8 * 8 * n_ckpt / max(args.steps, 1)),
                'note': 'the chain engines keep all state on the device; the host reads the statistics table at checkpoints'},
        'gpu_launches': int(_native.launch_count() - launches0),
        'checkpoints': n_ckpt,
        'per_size': {f'D={D}': {'chains': len(rows), 'median_best_loss': float(np.median([r[1] for r in rows])),
                                'acceptance_rate': float(sum(r[2] + r[4] for r in rows) / max(1.0, sum(r[3] + r[5] for r in rows)))}
                     for D, rows in sorted(by_d.items())},
        'clocks': clocks,
    }
    if rank == 0:
        emit_line(line)
    sweep.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit_line(line):
    """The ONE JSON line, on the process's real stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL's version banner, for one) write to file descriptor 1: keep stdout for the JSON line alone
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'multiset':
        run_multiset(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
