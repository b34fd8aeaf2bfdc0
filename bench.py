#!/usr/bin/env python
"""bench.py — FlowDec postfilter throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference algorithm on host CPU

Metric: 48 kHz audio-seconds generated per wall-second for `flowdec_75m` at NFE = 6
(midpoint, N = 3).  One "step" = one `FlowModel.enhance` pass over one batch of synthetic
clips: per GPU 32 clips x 2 s (BASELINE.json configs[1]); N GPUs = N independent shards, no
data-path collective (weak scaling).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
MACS_PER_FRAME = 5_861_842_944  # conv MACs per padded STFT frame column


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=2.0, help="clip length")
    ap.add_argument("--N", type=int, default=3)
    ap.add_argument("--solver", default="midpoint")
    ap.add_argument("--variant", default="75m")
    ap.add_argument("--max-batch", type=int, default=0, help="clips per backbone pass (0 = model default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3 / config4 / shard-check records")
    ap.add_argument("--streams", type=int, default=0, help="micro-batches in flight (0 = model default)")
    ap.add_argument("--max-frames", type=int, default=0, help="padded STFT frames per backbone pass (0 = model default)")
    ap.add_argument("--max-ctas", type=int, default=0, help="CTAs of the persistent conv kernels (0 = one per SM)")
    return ap.parse_args()


def nfe_of(N, solver):
    return N * (1 if solver == "euler" else 2)


def padded_frames(L):
    T = 1 + L // 384
    return T + (64 - T % 64) % 64


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"
    return 1400.0, "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("hbm_gbs", 6500.0))
    return 6500.0


# ------------------------------------------------------------------------------------------------
def pick_cpu_threads():
    """the oracle's convs are MKL-DNN bound; on many-core hosts fewer threads can be faster.
    Calibrate on one 256->256 3x3 conv and use the fastest of {all, 64, 32, 16} threads."""
    import torch.nn.functional as F
    n = os.cpu_count() or 1
    x = torch.randn(1, 256, 192, 32)
    w = torch.randn(256, 256, 3, 3)
    best, best_t = n, None
    for c in sorted({n, min(n, 64), min(n, 32), min(n, 16)}, reverse=True):
        torch.set_num_threads(c)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


_CPU_STATE = {}


def cpu_port_time(threads, seconds_clip=2.0, N=3, solver="midpoint"):
    """Times the CPU oracle (port of the reference algorithm) DIRECTLY on one clip of the workload's own length
    and solver setting: 1 clip x `seconds_clip` at (N, solver).  The reference processes clips
    independently and its cost is linear in the batch, so audio-s/s of the whole batch = seconds_clip / time."""
    from flowdec_b200.model import build_flowdec
    from flowdec_b200.util.synth import synth_state_dict, synth_waveforms
    from oracle import flowdec_oracle as O
    torch.set_num_threads(threads)
    if "sd" not in _CPU_STATE:
        _CPU_STATE["sd"] = synth_state_dict(build_flowdec("75m").state_dict(), seed=0)
        with torch.no_grad():      # page in MKL-DNN / the thread pool outside the timed sample
            O.enhance(_CPU_STATE["sd"], synth_waveforms(1, 12000, seed=1), N=1, solver="euler",
                      eps=torch.zeros(1, 1, 768, 64, dtype=torch.complex64))
    sd = _CPU_STATE["sd"]
    L = int(seconds_clip * SR)
    y = synth_waveforms(1, L, seed=1234)
    g = torch.Generator().manual_seed(4321)
    eps = torch.randn(1, 1, 768, padded_frames(L), dtype=torch.complex64, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.enhance(sd, y, N=N, solver=solver, eps=eps)
    dt = time.perf_counter() - t0
    nfe = nfe_of(N, solver)
    return dt, (f"oracle port, 1 clip x {seconds_clip:g} s ({padded_frames(L)} frames), {solver} N={N} (NFE {nfe}) timed "
                f"directly: {dt:.1f} s on {threads} threads; batch scaled linearly (clips are independent)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = pick_cpu_threads()
    times, sample = [], ""
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt, sample = cpu_port_time(threads, args.seconds, args.N, args.solver)
        # every pass is the same deterministic CPU work: under the time budget warm-up passes count as samples
        if i >= args.warmup or (time.perf_counter() - t_start) > 60:
            times.append(dt)
        if (time.perf_counter() - t_start) > 150:      # keep the whole run within a few minutes
            break
    dt = sum(times) / len(times)
    t_step = dt * args.batch * args.gpus               # one step = the whole batch, clip after clip
    value = args.seconds / dt
    line = {
        "impl": "reference", "metric": "48kHz audio-seconds/sec (RTF^-1) flowdec_75m NFE=6", "value": value,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    L = int(args.seconds * SR)
    return {"workload": f"flowdec_{args.variant} postfilter enhance(): {args.batch} clips x {args.seconds:g} s @48kHz per GPU, "
                        f"{args.solver} N={args.N} (NFE {nfe_of(args.N, args.solver)}), Tp={padded_frames(L)} frames",
            "per_gpu_batch": args.batch, "clip_seconds": args.seconds, "nfe": nfe_of(args.N, args.solver),
            "solver": args.solver, "sharding": f"dp{args.gpus} (clip batch, no data-path collective)",
            "l2": "working set (multi-GB activations per micro-batch) >> 126 MB L2; no explicit flush",
            "micro_batch": "16 clips (<= 4096 padded frames) per backbone pass, 2 passes in flight on separate CUDA streams (model.max_batch / overlap_streams)"}


def ndac_pipeline_record(model, args, dev, steps=3):
    """codes -> quantizer.from_codes -> DAC.decode -> FlowModel.enhance (demo.ipynb:104-109) on one GPU, with the
    share of each stage (CUDA events).  Synthetic NDAC-75-scale decoder: latent 1024, decoder_dim 1536, rates
    8*5*4*4 = 640 (75 Hz frames at 48 kHz), 10 codebooks; exact dims live in the (offline-unavailable) checkpoint."""
    from flowdec_b200.ndac import DAC
    from flowdec_b200.util.synth import synth_dac_state_dict
    rates, nq, latent, dim = (8, 5, 4, 4), 10, 1024, 1536
    dac = DAC(synth_dac_state_dict(latent, dim, rates, nq, seed=7), decoder_dim=dim, decoder_rates=rates,
              n_codebooks=nq, latent_dim=latent, sample_rate=SR).to(dev).eval()
    B = args.batch
    Tz = int(args.seconds * SR) // 640
    codes = torch.randint(0, 1024, (B, nq, Tz), generator=torch.Generator().manual_seed(3)).to(dev)
    import gc
    gc.collect()
    torch.cuda.empty_cache()      # the earlier legs leave tens of GB cached: start this one from a clean allocator
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    samples = []
    for it in range(steps + 3):
        ev[0].record()
        zq, _, _ = dac.quantizer.from_codes(codes)
        ev[1].record()
        xh = dac.decode(zq)
        ev[2].record()
        out = model.enhance(xh, N=args.N, solver=args.solver)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:                # eager call, graph capture and first replay of enhance() are warm-up
            samples.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
    assert torch.isfinite(out).all()
    ms = [sorted(s[i] for s in samples)[len(samples) // 2] for i in range(3)]      # median of the timed iterations
    model.reset_cache()
    return {"workload": f"{B} x {args.seconds:g} s: codes [B,{nq},{Tz}] -> from_codes -> decode (dim {dim}, rates {rates}) "
                        f"-> enhance NFE {nfe_of(args.N, args.solver)}; synthetic weights",
            "value": B * args.seconds / (sum(ms) * 1e-3), "unit": "audio-s/s", "from_codes_ms": ms[0],
            "decode_ms": ms[1], "enhance_ms": ms[2], "decode_share": ms[1] / sum(ms),
            "decode_audio_s_per_s": B * args.seconds / (ms[1] * 1e-3),
            # algorithmic HBM traffic of the decoder: 0.69 GB per audio-second (fd_dac_tc.cu row)
            "decode_roofline": {"bound": "hbm", "unit": "GB/s", "achieved": 0.69 * B * args.seconds / (ms[1] * 1e-3),
                                "peak": hbm_peak(), "frac": 0.69 * B * args.seconds / (ms[1] * 1e-3) / hbm_peak()}}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from flowdec_b200 import _lib, ops
    from flowdec_b200.model import build_flowdec
    from flowdec_b200.util.synth import synth_state_dict, synth_waveforms

    model = build_flowdec(args.variant)
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=0))
    model = model.to(dev)
    if args.max_batch:
        model.max_batch = args.max_batch
    if args.streams:
        model.overlap_streams = args.streams
    if args.max_frames:
        model.max_frames_per_pass = args.max_frames
    if args.max_ctas:
        model.backbone.max_ctas = args.max_ctas
    L = int(args.seconds * SR)
    B = args.batch
    nfe = nfe_of(args.N, args.solver)
    Tp = padded_frames(L)
    # clips are indexed globally so that every shard is the same regardless of world size
    y_host = synth_waveforms(B, L, seed=1234 + rank * B).pin_memory()
    y_dev = y_host.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first call eager + packs weights, second captures the CUDA graph) ----
    for _ in range(max(args.warmup, 3)):
        out = model.enhance(y_dev, N=args.N, solver=args.solver)
    torch.cuda.synchronize()

    # ---- kernel launches per step (eager count) ----
    model.use_cuda_graph = False
    n0 = _lib.LAUNCHES
    model.enhance(y_dev, N=args.N, solver=args.solver)
    torch.cuda.synchronize()
    launches_per_step = _lib.LAUNCHES - n0
    model.use_cuda_graph = True

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu_range = bool(os.environ.get("FD_NCU_RANGE"))   # `ncu --profile-from-start off` captures only this region
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        out = model.enhance(y_dev, N=args.N, solver=args.solver)
    e1.record()
    if ncu_range:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- end to end: pinned host input -> H2D -> enhance -> D2H ----
    for _ in range(1):
        model.enhance(y_host, N=args.N, solver=args.solver)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        out_host = model.enhance(y_host, N=args.N, solver=args.solver)   # returns on y's device (CPU): D2H inside
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    assert out_host.device.type == "cpu" and torch.isfinite(out_host).all()

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    audio_s = B * world * args.seconds * args.steps
    value = audio_s / (ms * 1e-3)
    value_e2e = audio_s / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 conv), one instrumented eager step on rank 0 ----
    roofline = None
    if rank == 0:
        model.use_cuda_graph = False
        lanes_saved, model.overlap_streams = model.overlap_streams, 1   # time the kernel alone on the SMs
        ops.PROFILE = []
        torch.cuda.synchronize()
        model.enhance(y_dev, N=args.N, solver=args.solver)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        model.use_cuda_graph = True
        model.overlap_streams = lanes_saved
        conv_ms = sum(a.elapsed_time(b) for a, b, _ in prof)
        conv_flops = sum(f for _, _, f in prof)
        peak, how = load_peaks()
        algo_flops_step = 2.0 * MACS_PER_FRAME * B * Tp * nfe
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_launch"]
            traffic_note = (f"STATIC ncu capture ({tj.get('source', 'profiles/')}), not measured in this run: "
                            f"dram read+write of the dominant launch shape ({tj['launch_shape']}); algorithmic "
                            f"{tj['algorithmic_bytes_per_launch']} B; tensor pipe active {tj['tensor_pipe_active_pct_of_elapsed']} % of elapsed")
        roofline = {"bound": "tensor", "kernel": "conv_halo_kernel / conv_igemm_kernel (tcgen05 implicit GEMM, GroupNorm+SiLU operand transform fused)", "achieved": achieved,
                    "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": how, "launches": len(prof), "kernel_ms_per_step": conv_ms,
                    "kernel_share_of_step": conv_ms / (ms / args.steps),
                    "algorithmic_tflop_per_step": algo_flops_step / 1e12,
                    "whole_step_tflops": algo_flops_step / (ms / args.steps * 1e-3) / 1e12,
                    "whole_step_frac": algo_flops_step / (ms / args.steps * 1e-3) / 1e12 / peak}

    # ---- secondary records (outside the headline timed region; own keys) ----
    def timed_config(m, batch, seconds, steps=2, N=None, solver=None):
        """same timing protocol (device-resident inputs, CUDA events, max over ranks) for another BASELINE config"""
        N, solver = N or args.N, solver or args.solver
        Lc = int(seconds * SR)
        yc = synth_waveforms(batch, Lc, seed=5000 + rank * batch).to(dev)
        for _ in range(3):
            m.enhance(yc, N=N, solver=solver)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            m.enhance(yc, N=N, solver=solver)
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt[0])
        m.reset_cache()
        return {"value": batch * world * seconds * steps / (t_ms * 1e-3), "unit": "audio-s/s",
                "ms_per_step": t_ms / steps, "steps": steps, "per_gpu_batch": batch, "clip_seconds": seconds}

    extras = {}
    if not args.no_extras:
        model.reset_cache()
        # BASELINE config 4: 256 x 4 s over 8 GPUs = 32 x 4 s per GPU (here: this many GPUs' worth of it)
        extras["config4"] = dict(timed_config(model, 32, 4.0),
                                 workload=f"flowdec_75m, {32 * world} x 4 s clips, NFE {nfe}, 32 per GPU on {world} GPU(s)")
        if world == 1:
            # BASELINE config 3: flowdec_25s (same backbone, its own sigma_y curve), 64 x 2 s
            m25 = build_flowdec("25s")
            m25.backbone, m25.feature_extractor = model.backbone, model.feature_extractor
            m25 = m25.to(dev)
            extras["config3"] = dict(timed_config(m25, 64, 2.0), workload=f"flowdec_25s, 64 x 2 s clips, NFE {nfe}, 1 GPU")
            del m25
            # BASELINE config 5: NFE sweep at batch 64 x 2 s (NFE 1 = Euler N=1; 2/4/8/16 = midpoint N=1/2/4/8)
            sweep = []
            for nfe5, (n5, s5) in ((1, (1, "euler")), (2, (1, "midpoint")), (4, (2, "midpoint")), (8, (4, "midpoint")),
                                   (16, (8, "midpoint"))):
                r5 = timed_config(model, 64, 2.0, steps=2, N=n5, solver=s5)
                sweep.append({"nfe": nfe5, "solver": s5, "N": n5, "value": r5["value"], "ms_per_step": r5["ms_per_step"]})
            extras["config5"] = {"workload": "flowdec_75m, 64 x 2 s clips, 1 GPU, NFE sweep", "unit": "audio-s/s",
                                 "sweep": sweep}
        if world == 1:
            # secondary line: the same workload with the backbone in tf32 precision (fp32 activations, kind::tf32 MMAs)
            model.set_precision("tf32")
            extras["tf32_mode"] = dict(timed_config(model, B, args.seconds),
                                       workload=f"headline workload with model.set_precision('tf32'): {B} x {args.seconds:g} s, "
                                                f"NFE {nfe}; parity 54 dB vs the reference goldens (bf16: 37 dB)")
            model.set_precision("bf16")
            extras["ndac_pipeline"] = ndac_pipeline_record(model, args, dev)
        if dist is not None:
            from flowdec_b200 import parallel
            ok = parallel.verify_sharding(model, dist, L, args.N, args.solver)
            extras["shard_bitwise_equal"] = ok
            extras["shard_check"] = (f"{2 * world} clips x {args.seconds:g} s: NCCL scatter from rank 0 -> enhance per rank "
                                     f"(noise seeded by global clip index) -> NCCL all_gather, compared bit for bit "
                                     f"with the same batch enhanced on rank 0 alone")
            model.reset_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = pick_cpu_threads()
        dt, sample = cpu_port_time(threads, args.seconds, args.N, args.solver)
        cpu_baseline = {"value": args.seconds / dt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                        "sample": sample}

    if rank == 0:
        line = {
            "metric": "48kHz audio-seconds/sec (RTF^-1) flowdec_75m NFE=6", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 (tensor-core operands, fp32 accumulate; STFT/ODE/GroupNorm statistics fp32)",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": value_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * L * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        line.update(extras)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
