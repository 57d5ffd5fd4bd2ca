#!/usr/bin/env python
"""bench.py — images/sec of one LSNet training step (R50-FPN, 800x1333 -> padded 800x1344, batch 4 per GPU, bf16),
BASELINE.json's metric on BASELINE.json's configs[1], on N GPUs of one node (one process per GPU, NCCL).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference path on the host cores

A "step" = forward + cross-IOU/focal loss + backward + grad-clip + SGD(momentum) on one synthetic batch.
``value``  : inputs already resident in HBM (K steps, CUDA events, barrier + synchronize on both sides, max over ranks)
``e2e``    : the same K steps through the public API with HOST (pinned) image tensors: H2D copy of the batch and a
             D2H read of the loss inside the timed region, every step.
"""
import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm mixes two OpenMP runtimes (torch's bundled libgomp and the system one behind oracle/dcn_ref.c); with
# active spinning they fight over the cores, so make idle workers sleep.  Must be set before either is loaded.
os.environ.setdefault('OMP_WAIT_POLICY', 'PASSIVE')
os.environ.setdefault('GOMP_SPINCOUNT', '0')

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'images/sec training step, LSNet R50-FPN 800x1333'
WORKLOAD = 'LSNet-bbox R50-FPN 800x1333 (padded 800x1344) bf16, batch 4/GPU, synthetic COCO-shaped'
IMG_HW = (800, 1333)
BATCH = 4
# --config: BASELINE.json configs[1] (the default, the configuration the metric is quoted on) and configs[2..4]
WORKLOADS = {
    'bbox_r50': dict(workload=WORKLOAD, multiscale=None),
    'bbox_x101dcn_ms': dict(workload='LSNet-bbox X-101-64x4d-DCN (DCNv2 groups=64 in c3-c5, with_cp), multi-scale short side '
                                     '480-960 / long side <= 1333, bf16, batch 4/GPU, synthetic COCO-shaped', multiscale=(480, 960)),
    'segm_r50': dict(workload='LSNet-seg R50-FPN, 36 contour landmarks, 800x1333 (padded 800x1344) bf16, batch 4/GPU, '
                              'synthetic COCO-shaped', multiscale=None),
    'pose_x101dcn': dict(workload='LSNet-pose X-101-64x4d-DCN, 17 keypoints + boxes (pose_bbox), 800x1333 (padded 800x1344) bf16, '
                                  'batch 4/GPU, synthetic COCO-shaped', multiscale=None),
}
# kernel classes of lsnet_timing_collect (lsnet_internal.h): name -> what bounds it ('tensor': work = FLOPs, 'hbm': bytes)
CLASS_NAMES = ['gemm_kmajor(tcgen05 GEMM/implicit conv)', 'gemm_mnmajor(tcgen05 weight grad)', 'dcn_im2col(gather)',
               'dcn_col2im(scatter)', 'dcn_fused_fwd(gather->smem->tcgen05)', 'dcn_fused_wgrad(gather->smem->tcgen05)',
               'dcn_fused_bwd_data(tcgen05->scatter)',
               # gemm_kmajor launches whose FLOPs / algorithmic bytes is below the ridge (~200 FLOP/B = measured tensor peak /
               # measured HBM peak): the trunk's K <= 256 1x1 convs and other thin GEMMs; accounted in bytes
               'gemm_kmajor_hbm(tcgen05 GEMM, intensity below the ridge)']
TENSOR_BOUND = ('gemm_kmajor(', 'gemm_mnmajor', 'dcn_fused')


def usable_cores():
    """Host cores this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def cpu_threads():
    # beyond ~32 threads the CPU path (many small convs / per-image target code) stops scaling and starts thrashing
    return max(1, min(usable_cores(), 32))


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
    """SM clock / throttle reasons every 200 ms during the timed region (B200_PROFILING.md recipe).  Sampled in-process
    through NVML (nvidia_ml_py): an `nvidia-smi -lms` child takes the driver's global lock for every query, which shows
    up as multi-ms stalls in the e2e loop (one synchronisation per step); falls back to nvidia-smi without NVML."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.stop_flag, self.thread = None, threading.Event(), None
        self.sm, self.smax, self.reasons = [], None, set()

    def _nvml_loop(self):
        n, h = self.nvml
        bits = {}
        for name, attr in zip(self.NAMES, ['HwSlowdown', 'HwThermalSlowdown', 'SwThermalSlowdown', 'SwPowerCap']):
            for pre in ('nvmlClocksEventReason', 'nvmlClocksThrottleReason'):
                if hasattr(n, pre + attr):
                    bits[name] = getattr(n, pre + attr)
                    break
        get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(n, 'nvmlDeviceGetCurrentClocksThrottleReasons', None)
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(h))
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            h = n.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=self.smax, reasons=sorted(self.reasons),
                        samples=len(sm), source='nvml')
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
                for n, v in zip(self.NAMES, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm), source='nvidia-smi')


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step_factory(threads):
    """The reference path restated for the CPU (oracle/): one training step = LSDetector fwd + loss + bwd + clip +
    SGD on a bounded sample.  Returns (step_fn, sample_description, images_per_step_equivalent)."""
    from oracle import init as oinit
    from oracle import lsnet_oracle as O
    from lsnet_b200.data import synthetic_batch
    torch.set_num_threads(threads)
    os.environ['OMP_NUM_THREADS'] = str(threads)
    try:
        ctypes.CDLL('libgomp.so.1').omp_set_num_threads(threads)   # the C restatement's runtime
    except OSError:
        pass
    sd = oinit.make_state_dict('bbox', seed=0)
    keys = O.trainable_keys(sd)
    mom = {}

    def run(hw, step):
        b = synthetic_batch(step, batch=1, img_hw=hw)
        for k in keys:
            sd[k].requires_grad_(True)
            sd[k].grad = None
        losses = O.detector_losses(sd, b['img'], b['gt_bboxes'], b['gt_labels'], b['img_metas'], task='bbox',
                                   gt_extremes=b['gt_extremes'])
        total, _ = O.parse_losses(losses)
        total.backward()
        grads = {k: sd[k].grad for k in keys if sd[k].grad is not None}
        for k in keys:
            sd[k].requires_grad_(False)
        O.sgd_step(sd, grads, mom)
        return float(total)
    return run


def bench_config(world, name='bbox_r50'):
    """The workload description both arms print verbatim (the driver compares the two `config` objects)."""
    return dict(workload=WORKLOADS[name]['workload'], global_batch=BATCH * world, parallelism=f'dp{world}',
                l2_policy='inputs+activations per step (>1 GB) far exceed the 126 MB L2; 4 rotating batches',
                optimizer='SGD lr0.01 m0.9 wd1e-4, grad-clip 35, fp32 master weights')


def run_reference(args, rank, world):
    """CPU arm: the reference path (oracle port, fp32) on the host cores, ALWAYS at the full 800x1333 resolution --
    one image per step is the bounded sample; the image is never shrunk, so the number means the same on every box."""
    if rank != 0:
        return
    threads = cpu_threads()
    run = cpu_reference_step_factory(threads)
    sample = '1 image 800x1333 per step (full resolution), all host threads'
    for w in range(args.warmup):
        run(IMG_HW, w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        run(IMG_HW, args.warmup + s)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = dict(metric=METRIC, value=value, unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000 * dt / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference', config=bench_config(world),
                step_mode='reference path on the host cores (oracle port, fp32), rank 0 only',
                cpu_baseline=dict(value=value, unit='images/s', cores=threads, kind='port', sample=sample),
                e2e=dict(value=value, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def measured_traffic(kernel_class):
    """DRAM bytes per launch of a kernel class inside the step (dram__bytes_read.sum + dram__bytes_write.sum), from the
    committed ncu pass over one whole step (profiles/r02_step_traffic.json, tools/step_traffic.py).  None when the
    profile does not cover the class."""
    p = os.path.join(ROOT, 'profiles', 'r02_step_traffic.json')
    try:
        d = json.load(open(p))['classes']
        key = kernel_class.split('(')[0]
        return float(d[key]['dram_bytes_per_launch']) if key in d else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch.distributed as dist
    from lsnet_b200 import lib as L
    from lsnet_b200.data import MODEL_CFG, TASK_OF, synthetic_batch, to_device
    from lsnet_b200.train import Trainer
    cfg_name = args.config
    task, ms = TASK_OF[cfg_name], WORKLOADS[cfg_name]['multiscale']
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    nb = 4
    # multi-scale: every image draws its own size, the canvas of a batch is rounded up to a multiple of 128 so that the
    # CUDA-graph cache sees a handful of shape buckets (the valid extents per image stay exact: pad_shape)
    host = [synthetic_batch(s, rank, BATCH, IMG_HW, pin=True, task=task, multiscale=ms, canvas_multiple=128 if ms else None)
            for s in range(nb)]
    resident = [to_device(b, dev) for b in host]
    mode = 'cuda-graph (fwd+loss+bwd captured; flat all-reduce + clip + SGD eager)'
    if args.eager:
        tr = Trainer(MODEL_CFG[cfg_name], device=dev, distributed=distributed)
        mode = 'eager (torch DDP)'
    else:
        from lsnet_b200.train import GraphTrainer
        tr = GraphTrainer(MODEL_CFG[cfg_name], host[0], device=dev, distributed=distributed, kernel_timing=True)
    torch.cuda.synchronize()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        # A cyclic-GC pass over this process's heap (modules, captured graphs, autograd records) takes 40-150 ms; in the
        # e2e loop (one host synchronisation per step) a single automatic collection showed up as one 44 / 165 ms step among
        # twenty 22.8 ms ones.  Collect before the region and keep the collector off inside it (as a training loop would).
        gc.collect()
        gc.disable()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            for s in range(steps):
                fn(s)
        finally:
            gc.enable()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms)

    for w in range(max(args.warmup, 3, nb if ms else 0)):     # multi-scale: every shape bucket captured before timing
        tr.step(resident[w % nb])
    lib = L.load()
    lib.lsnet_launch_count.restype = ctypes.c_ulonglong
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # ---- value: inputs resident in HBM; kernel-class timing (CUDA events on the launch stream) enabled ----
    graph_mode = not args.eager
    if not graph_mode:
        lib.lsnet_timing_reset()
        lib.lsnet_timing_enable(1)
    l0 = L.launch_count()
    ms = timed(lambda s: tr.step(resident[s % nb]), args.steps)
    launches = (tr.launches_per_step * args.steps) if graph_mode else (L.launch_count() - l0)
    lib.lsnet_timing_enable(0)
    # eager: events of all K steps.  graph: one extra replay of the INSTRUMENTED capture of the same step right after
    # the timed region (external event-record nodes; the timed graph itself carries no instrumentation)
    timed_steps = 1 if graph_mode else args.steps
    serial_ms = None
    if graph_mode:
        tr.replay_instrumented()                  # warm (first replay of this graph)
        serial_ms = tr.replay_instrumented()
        torch.cuda.synchronize()
    classes = {}
    for cls, name in enumerate(CLASS_NAMES):
        tms, n, work = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        lib.lsnet_timing_collect(cls, ctypes.byref(tms), ctypes.byref(n), ctypes.byref(work))
        if n.value:
            classes[name] = dict(ms=tms.value, launches=n.value, work=work.value)
    if not graph_mode:
        lib.lsnet_timing_reset()
    # ---- e2e: host (pinned) inputs -> H2D every step, loss read back every step ----
    if args.e2e_input == 'u8' and graph_mode:
        mscale = WORKLOADS[cfg_name]['multiscale']
        host = [synthetic_batch(s, rank, BATCH, IMG_HW, pin=True, task=task, multiscale=mscale,
                                canvas_multiple=128 if mscale else None, u8=True) for s in range(nb)]
    h2d = sum(b['img'].numel() * b['img'].element_size() for b in host) // nb

    e2e_wall = []

    def e2e_step(s):
        # graph mode: the step takes the pinned HOST batch and prefetches the next one on a copy stream while it computes
        # (every step's inputs cross PCIe inside the timed region, overlapped with the previous step); eager: .to(device)
        t0 = time.perf_counter()
        if graph_mode:
            loss, _ = tr.step(host[s % nb], next_batch=host[(s + 1) % nb])
        else:
            loss, _ = tr.step(to_device(host[s % nb], dev))
        loss.item()
        e2e_wall.append(1e3 * (time.perf_counter() - t0))
    for w in range(3):           # the host-input path has its own first-use work (pinned staging sets): warm it up too
        e2e_step(w)
    e2e_wall.clear()
    ms_e2e = timed(e2e_step, args.steps)
    wall = sorted(e2e_wall[-args.steps:])
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    imgs = BATCH * world * args.steps
    value, e2e = imgs / (ms / 1e3), imgs / (ms_e2e / 1e3)
    # dominant kernel class = the one with the largest share of the timed region
    dom = max(classes, key=lambda k: classes[k]['ms'])
    c = classes[dom]
    per_launch_ms = c['ms'] / max(1, c['launches'])
    if any(t in dom for t in TENSOR_BOUND):
        achieved = c['work'] / (c['ms'] / 1e3) / 1e12 if c['ms'] > 0 else 0.0
        roof = dict(bound='tensor', kernel=dom, achieved=achieved, peak=peaks['tf_sustained'], unit='TFLOP/s',
                    frac=achieved / peaks['tf_sustained'], traffic=measured_traffic(dom))
    else:
        achieved = c['work'] / (c['ms'] / 1e3) / 1e9 if c['ms'] > 0 else 0.0
        roof = dict(bound='hbm', kernel=dom, achieved=achieved, peak=peaks['hbm'], unit='GB/s',
                    frac=achieved / peaks['hbm'], traffic=measured_traffic(dom))
    roof.update(peak_source=peaks['src'] + (' sustained' if any(t in dom for t in TENSOR_BOUND) else ''),
                # graph mode: kernel classes are timed in a SERIALISED instrumented replay of the same step (stream
                # parallelism off, every kernel alone on the GPU, like an ncu launch list); shares are of that replay
                share_of_step=c['ms'] / (serial_ms if serial_ms else ms * timed_steps / args.steps),
                serialized_step_ms=serial_ms, avg_launch_ms=per_launch_ms,
                launches_timed=c['launches'], timed_steps=timed_steps,
                classes={k: dict(ms_per_step=v['ms'] / timed_steps, launches_per_step=v['launches'] / timed_steps,
                                 achieved=(v['work'] / (v['ms'] / 1e3) / (1e12 if any(t in k for t in TENSOR_BOUND) else 1e9))
                                 if v['ms'] > 0 else 0.0,
                                 unit='TFLOP/s' if any(t in k for t in TENSOR_BOUND) else 'GB/s') for k, v in classes.items()})
    cpu = None
    if world == 1 and not args.no_cpu_baseline and cfg_name == 'bbox_r50':
        try:
            threads = cpu_threads()
            run = cpu_reference_step_factory(threads)
            run(IMG_HW, 0)                                   # warm-up (first-touch, thread pools)
            t0 = time.perf_counter(); n = 0
            while n < 3 and (n == 0 or time.perf_counter() - t0 < 18.0):
                run(IMG_HW, 1 + n); n += 1
            dt = time.perf_counter() - t0
            cpu = dict(value=n / dt, unit='images/s', cores=threads, kind='port',
                       sample=f'{n} training steps on 1 image 800x1333 each (oracle port of the reference path, fp32, '
                              'full resolution, all host threads)')
        except Exception as e:   # the baseline must never take the GPU number down with it
            cpu = dict(value=None, unit='images/s', cores=cpu_threads(), kind='port', sample=f'failed: {e!r}')
    line = dict(metric=METRIC, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                data='synthetic', config=bench_config(world, cfg_name), step_mode=mode,
                shapes=sorted({tuple(b['img'].shape[-2:]) for b in resident}),
                e2e=dict(value=e2e, unit='images/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                         ms_per_step=ms_e2e / args.steps,
                         input=('uint8 HWC bytes, normalised + padded on the GPU (lsnet_image_prep_u8) inside the step'
                                if host[0]['img'].dtype == torch.uint8 else 'normalised fp32 NCHW batch'),
                         # host wall clock of the individual steps (each ends with the loss read-back): spread of the region
                         step_wall_ms=dict(min=wall[0], median=wall[len(wall) // 2], max=wall[-1])),
                gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu, clocks=clk)
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--config', default='bbox_r50', choices=sorted(WORKLOADS),
                    help="bbox_r50 = BASELINE.json configs[1] (default, the metric's configuration); the others are "
                         'configs[2..4]')
    ap.add_argument('--e2e-input', default='u8', choices=['f32', 'u8'],
                    help='host image format of the e2e loop: f32 = normalised float batch (51.6 MB / step), u8 = decoded '
                         'bytes, normalised + padded on the GPU by lsnet_image_prep_u8 inside the step (12.9 MB / step)')
    ap.add_argument('--eager', action='store_true', help='per-op eager step with torch DDP instead of the CUDA graph')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        if args.config != 'bbox_r50':
            raise SystemExit('--impl reference is the CPU arm of the default configuration (bbox_r50) only')
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
