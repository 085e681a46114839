#!/usr/bin/env python
"""bench.py -- frames/s of the deephar forward hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm, see below)

Workload (config.workload): ReceptionNet 2-D pose, the model of BASELINE.json configs[1]
(`reception.build((256,256,3), 16, dim=2, num_blocks=8, num_context_per_joint=2, ksize=(5,5))`,
exp/mpii/eval_mpii_singleperson.py:42-49) on the headline batch: 32 clips x 16 frames of
256x256x3 = 512 frames per GPU per step (TimeDistributed folds clips into frames), synthetic
uniform[-1,1] frames and seeded synthetic weights (no datasets / checkpoints offline).

  value  : frames/s, inputs resident in HBM (512 frames = 403 MB > L2, so no L2 flush is
           needed between iterations), CUDA-event timed, max over ranks.
  e2e    : frames/s through Model.predict() with pinned HOST input, H2D + D2H inside the
           timed region.
  --impl reference : the reference's Keras/TF forward cannot run here (no tensorflow/keras in
           the image); the arm times the CPU port of the same graph
           (oracle/, torch-CPU fp32, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_KW = dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=8, ksize=(5, 5),
                concat_pose_confidence=False)
CLIPS, FRAMES = 32, 16
METRIC = 'frames/sec (256x256, 16-frame clips, b32)'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_port_frames_per_sec(n_frames, batch, warmup=1):
    """CPU port (oracle, torch-CPU fp32/oneDNN, all threads) of the same model on a bounded sample."""
    import torch
    from deephar_b200 import reception
    from oracle import ops_torch, synth
    from oracle import reception as oracle_reception
    m = reception.build((256, 256, 3), **MODEL_KW).init_synthetic_weights(1234)
    table = m.get_weights()
    x = synth.synth_frames(batch, seed=0)
    for _ in range(warmup):
        oracle_reception.forward(ops_torch, table, x, **MODEL_KW)
    t0 = time.perf_counter()
    done = 0
    while done < n_frames:
        oracle_reception.forward(ops_torch, table, x, **MODEL_KW)
        done += batch
    dt = time.perf_counter() - t0
    return done / dt, torch.get_num_threads(), done, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = 8           # frames per "step" of the bounded CPU sample
    for _ in range(max(args.warmup, 1) - 1):
        cpu_port_frames_per_sec(per_step, 4, warmup=0)
    fps, cores, done, dt = cpu_port_frames_per_sec(per_step * args.steps, 4, warmup=1)
    sample = '%d frames of the 512-frame step (batches of 4), torch-CPU fp32 port of the Keras graph' % done
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        # same workload as the product arm; each step is a bounded sample of it (see cpu_baseline.sample)
        'config': {'workload': 'reception2d_8blk_k5_j16 (BASELINE configs[1] model) x 32 clips x 16 frames',
                   'global_batch_frames': CLIPS * FRAMES, 'sample_frames_per_step': per_step,
                   'implementation': 'CPU port of the Keras graph (oracle/, torch-CPU fp32, all host threads); '
                                     'keras 2.1.4 / tensorflow 1.6 are not installable in this image'},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def softargmax_microbench(torch, model, peaks):
    """dh_softargmax2d_ctx_f32 on 4096 frames of (32,32,48) = 805 MB (6x L2): achieved HBM GB/s."""
    import ctypes as C
    from deephar_b200 import _ffi
    n = 4096
    g = torch.Generator(device='cuda').manual_seed(0)
    h = torch.randn(n, 32, 32, 48, device='cuda', generator=g) * 3.0
    pose = torch.empty(n, 16, 2, device='cuda')
    vis = torch.empty(n, 16, 1, device='cuda')
    hv = _ffi.dh_view(h.data_ptr(), n, 32, 32, 48, 48)
    lib, ctx = _ffi.lib(), model._ctx.handle
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        _ffi.check(lib.dh_softargmax2d_ctx_f32(ctx, C.byref(hv), 16, 2, C.c_float(0.8), pose.data_ptr(),
                                               vis.data_ptr(), st), 'softargmax')
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_ = n * (32 * 32 * 48 * 4 + 16 * 3 * 4)
    gbs = bytes_ / ms / 1e6
    return {'kernel': 'softargmax2d_ctx (32x32x48 maps, %d frames, %.0f MB > L2)' % (n, bytes_ / 1e6),
            'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
            'frac': gbs / peaks['hbm_gbs'], 'us_per_launch': ms * 1000.0, 'peak_source': peaks['source']}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture of the dominant kernel (committed
# under profiles/), per frame; the capture is a 128-frame launch, the bench launch is micro_batch frames.
NCU_TRAFFIC = {
    'sepconv 32x32x576->32x32x576 k5x5': {
        'bytes_per_frame': (855.118336e6 + 270.342912e6) / 128,
        'source': 'profiles/r1_septma_final.ncu-rep (128-frame launch: 855.1 MB read + 270.3 MB written), scaled per frame'},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--micro-batch', type=int, default=256, help='frames per forward call')
    ap.add_argument('--workload', default='reception2d', choices=['reception2d', 'spnet_penn', 'spnet_ntu'],
                    help='reception2d = BASELINE configs[1] model (headline); spnet_* = configs[3]/[4] models')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--precision', type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    peaks = measured_peaks()
    n_frames = CLIPS * FRAMES
    mb = args.micro_batch
    assert n_frames % mb == 0
    if args.workload == 'reception2d':
        from deephar_b200 import reception
        model = reception.build((256, 256, 3), **MODEL_KW).init_synthetic_weights(1234)
        in_shape, step_items, wl_name = (n_frames, 256, 256, 3), mb, \
            'reception2d_8blk_k5_j16 (BASELINE configs[1] model) x 32 clips x 16 frames'
    else:
        from deephar_b200 import spnet
        from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d
        if args.workload == 'spnet_penn':
            cfg = ModelConfig((FRAMES, 256, 256, 3), pa16j2d, num_actions=[15], num_pyramids=6,
                              action_pyramids=[5, 6], num_levels=4, pose_replica=True, num_pose_features=160,
                              num_visual_features=160)
            wl_name = 'spnet PennAction multitask (BASELINE configs[3] model) x 32 clips x 16 frames'
        else:
            cfg = ModelConfig((FRAMES, 256, 256, 3), pa17j3d, num_actions=[60], num_pyramids=2,
                              action_pyramids=[1, 2], num_levels=4, num_pose_features=192, num_visual_features=192)
            wl_name = 'spnet NTU 3D multitask (BASELINE configs[4] model) x 32 clips x 16 frames'
        model = spnet.build(cfg).init_synthetic_weights(1234)
        assert mb % FRAMES == 0
        in_shape, step_items = (CLIPS, FRAMES, 256, 256, 3), mb // FRAMES
    model.precision = args.precision

    # synthetic frames, uniform [-1,1], pinned host memory (e2e source) + a device copy (value)
    gen = torch.Generator().manual_seed(rank)
    x_host = torch.empty(*in_shape, dtype=torch.float32).pin_memory()
    x_host.uniform_(-1.0, 1.0, generator=gen)
    x_dev = x_host.cuda()

    def step_device():
        last = None
        for i in range(0, in_shape[0], step_items):
            last = model.forward_device(x_dev[i:i + step_items])
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_outputs(outs):
        # the one exchange step of the data-parallel path: all-gather of the final outputs
        # -- reception: last block's (pose, vis); SPNet: last action probabilities (B_local, n_act)
        from deephar_b200.dist import gather_outputs as dh_gather
        local_out = (torch.cat([outs[-2], outs[-1]], dim=-1) if args.workload == 'reception2d' else outs[-1])
        return dh_gather(local_out.contiguous(), world)

    for _ in range(args.warmup):
        gather_outputs(step_device())
    barrier()
    model._ctx.launch_count(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        gather_outputs(step_device())
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = model._ctx.launch_count(reset=True)
    clocks = sampler.stop()
    t = torch.tensor([ms_total], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n_frames / (ms_step / 1000.0)

    # ---- e2e: public API, pinned host input, H2D + D2H inside the timed region ----
    x_np = x_host.numpy()
    for _ in range(2):
        model.predict(x_np[:2 * step_items], batch_size=step_items)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs = model.predict(x_np, batch_size=step_items)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_frames * args.steps / float(t.item())
    d2h = sum(int(np.prod(o.shape)) * 4 for o in outs)

    # ---- per-kernel profile (CUDA events around every launch of one extra step) ----
    prof = model.profile(x_dev[:step_items])
    conv_flops = model.conv_flops_per_frame()
    top = max(prof.values(), key=lambda r: r['ms'])
    total_ms = sum(r['ms'] for r in prof.values())
    tf = top['flops'] / (top['ms'] / top['launches'] / 1000.0) / 1e12 if top['flops'] else 0.0
    roofline = {
        'kernel': top['label'], 'bound': 'tensor', 'achieved': tf, 'peak': peaks['bf16_tflops_sustained'],
        'unit': 'TFLOP/s', 'frac': tf / peaks['bf16_tflops_sustained'],
        'traffic': NCU_TRAFFIC.get(top['label'], {}).get('bytes_per_frame', 0) * step_items * (1 if args.workload == 'reception2d' else FRAMES) or None,
        'traffic_source': NCU_TRAFFIC.get(top['label'], {}).get('source'),
        'share_of_step': top['ms'] / total_ms, 'us_per_launch': 1000.0 * top['ms'] / top['launches'],
        'peak_source': peaks['source'] + ' bf16 dense (sustained); kernel math: ' + model.math_mode(),
        'whole_forward_tflops': conv_flops * n_frames / (ms_step / 1000.0) / 1e12,
    }
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': model.math_mode(), 'data': 'synthetic',
        'config': {'workload': wl_name,
                   'global_batch_frames': world * n_frames, 'frames_per_gpu': n_frames, 'micro_batch': mb,
                   'parallelism': 'dp%d' % world,
                   'l2': 'inputs 403 MB per step > 126 MB L2; no flush needed'},
        'clocks': clocks,
        'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': n_frames * 256 * 256 * 3 * 4,
                'd2h_bytes_per_step': d2h},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'softargmax': softargmax_microbench(torch, model, peaks),
        'kernel_profile': sorted(([r['label'], round(r['ms'], 3), r['launches']] for r in prof.values()),
                                 key=lambda r: -r[1])[:8],
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, cores, done, dtc = cpu_port_frames_per_sec(24, 4)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d frames (batches of 4) of the same model, torch-CPU fp32 port '
                                          'of the Keras graph, %.1f s' % (done, dtc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception as e:      # the JSON line is already out; never fail the run on teardown
            print('teardown: %r' % (e,), file=sys.stderr)


if __name__ == '__main__':
    main()
