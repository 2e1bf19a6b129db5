#!/usr/bin/env python
"""bench.py -- the measurement contract for the hot path (BASELINE.json configs[1]).

Workload: stream-mHC layer microbenchmark, n = 4 streams, C = 512, T = 2^20 tokens per GPU, bf16,
forward + backward, 20 Sinkhorn iterations, synthetic inputs, random-init parameters
(SURVEY.md section 8(d), config 2).  One "step" = one forward + one backward over the T tokens.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--tokens T] [--impl reference]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU).  Tokens are sharded
across ranks with no data-path collective ("weak": T tokens per GPU); like a DDP step, the small
parameter gradients are all-reduced over NCCL after the backward.  Rank 0 prints ONE JSON line.

value        tokens/s (whole job) with x, dy resident in HBM, CUDA-event timed, max over ranks
e2e          the same metric through the host-buffer entry (stream_mhc_fwd_bwd_host): pinned host x, dy
             -> H2D -> kernels -> D2H of y, dx and the parameter gradients, all inside the timed region
roofline     dominant kernel (the fused single-pass backward, dx + every parameter gradient): algorithmic
             12288 B/token over its mean CUDA-event duration across the timed steps (hvs_mhc_stream_profile
             hooks: events on the launching stream, read back after the timed region), against
             MEASURED_PEAKS.json.  The training forward also writes, and the backward reads, 112 B/token of
             saved statistics (1.1 % on top of the 20480 B/token; not counted in the algorithmic figure)
cpu_baseline oracle/ (a port of the reference's PyTorch arithmetic) timed on the host cores, rank 0,
             N = 1 only, bounded sample of the same workload
k2           BASELINE.md's second reading of configs[1]: ManifoldHyperConnection(512, expansion_rate=4) on [2^20, 512]
             bf16 (tensor-core bound): TFLOP/s of the fused tcgen05 path and of the same module's library path
detect       decode + two-stage NMS at batch 64 / 640x640 grids (SURVEY D18 worst case and objectness bias -4)
hybrid_vision  the other half of BASELINE.json's metric ("hybrid_vision img/s at 1/2/4/8 B200"): configs[2] batch-64
             inference sharded 64/N with no collective, configs[3] bf16 DDP training at 16 images / GPU (gradient
             all-reduce over NCCL), configs[4] streaming batch-1 p50 / p99 under a CUDA graph; configs[0] (the CPU
             forward) is its cpu_baseline.  --skip-hybrid / --skip-extras leave these legs out
adaptive_sinkhorn  the same K1 step with the OPT-IN HVS_MHC_ADAPTIVE_ITERS flag (Sinkhorn loops stop at convergence to 2^-20;
             results within 2e-6 of the full run) -- reported next to the headline, which runs all 20 iterations
             hybrid_vision.training runs the whole optimisation step (DDP all-reduces included) as ONE CUDA graph replay
             (--train-eager for the eager step); all 76 mHC layers train on this library's forward and backward kernels
clocks       SM clock and throttle reasons sampled through NVML every 2 ms inside the timed region
--impl reference   times that CPU implementation alone (the reference is pure PyTorch; its own modules
             cannot travel to the GPU box)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_STREAMS, CHANNELS, LOGITS = 4, 512, 24
FWD_BYTES, BWD_BYTES = 8192, 12288                   # algorithmic bytes per token
METRIC, UNIT = "mhc_layer_fwd_bwd_tokens_per_s", "tokens/s"
FALLBACK_HBM_GBS = 6650.0                            # B200_PROFILING.md fallback


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def load_traffic():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs: NVML polled every 2 ms from a thread
    (the region can be a few tens of ms, too short for nvidia-smi's 100 ms loop), nvidia-smi as the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.nvml, self.handle, self.thread, self.running = None, None, None, False
        self.sm, self.max_sm, self.reasons = [], None, set()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                                            # CUDA index -> physical GPU through the UUID
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def _poll(self):
        n = self.nvml
        masks = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                 (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                 (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                 (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.handle))
                for m, name in masks:
                    if r & m:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_sm = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml, 2 ms poll over the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(self.NAMES, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def cpu_reference_run(tokens: int, steps: int, warmup: int):
    """Oracle forward+backward (fp32, all host threads) on `tokens` tokens; returns tokens/s and meta."""
    import torch
    from oracle import mhc_ref
    torch.manual_seed(0)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(tokens, N_STREAMS, CHANNELS, generator=g).to(torch.bfloat16)
    dy = torch.randn(tokens, N_STREAMS, CHANNELS, generator=g).to(torch.bfloat16)
    phi = torch.randn(N_STREAMS * CHANNELS, LOGITS, generator=g) * 0.02
    bias, alpha, scale = torch.zeros(LOGITS), torch.full((3,), 0.01), torch.ones(N_STREAMS * CHANNELS)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mhc_ref.stream_mhc_backward(x, dy, phi, bias, alpha, scale)      # forward + autograd backward
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return tokens / dt, dt, torch.get_num_threads()


def cpu_hybrid_baseline():
    """BASELINE configs[0]: whole-model CPU forward (+ decode + two-stage NMS), batch 1 at 640x640, fp32, all host
    threads -- the oracle port of the reference path (oracle/hybrid_ref.py)."""
    import torch
    from oracle import hybrid_ref
    torch.set_num_threads(os.cpu_count() or 1)
    ips, dt, threads = hybrid_ref.cpu_inference_img_per_s(batch=1, image=640, steps=2, warmup=1)
    return {"value": ips, "unit": "img/s", "cores": threads, "kind": "port", "s_per_image": dt,
            "sample": "hybrid_vision forward + decode + two-stage NMS, batch 1, 640x640, fp32 (BASELINE configs[0]); host wiring of "
                      "hvs_b200/hybrid_vision.py with oracle/ CPU leaves, mean of 2 after 1 warm-up"}


def cpu_k2_baseline(tokens: int = 8192, dim: int = 512, expansion: int = 4):
    """BASELINE.md's K2 reading of configs[1] on the host cores: the reference-literal module's eval forward (oracle port,
    fp32, constrained matrices recomputed per call like the reference) on a bounded token sample."""
    import torch
    from oracle import mhc_ref
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    h = dim * expansion
    p = {"H_pre_raw": torch.randn(dim, h), "H_post_raw": torch.randn(h, dim), "H_res_raw": torch.randn(dim, dim),
         "mlp.0.weight": torch.randn(2 * h, h) * 0.02, "mlp.0.bias": torch.zeros(2 * h), "mlp.3.weight": torch.randn(h, 2 * h) * 0.02,
         "mlp.3.bias": torch.zeros(h), "norm_pre.weight": torch.ones(dim), "norm_pre.bias": torch.zeros(dim),
         "norm_post.weight": torch.ones(dim), "norm_post.bias": torch.zeros(dim)}
    x = torch.randn(tokens, dim)
    with torch.no_grad():
        mhc_ref.mhc_module_forward(x, p)
        t0 = time.perf_counter()
        for _ in range(2):
            mhc_ref.mhc_module_forward(x, p)
        dt = (time.perf_counter() - t0) / 2
    flop = 2.0 * (2 * dim * h + 4 * h * h + dim * dim) * tokens
    return {"value": tokens / dt, "unit": "tokens/s", "tflops": flop / dt / 1e12, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{tokens} tokens, ManifoldHyperConnection({dim}, {expansion}) eval forward, oracle/mhc_ref.py (torch fp32 CPU), mean of 2 after 1 warm-up"}


def cpu_detect_baseline():
    import torch
    from oracle import detect_ref
    g = torch.Generator().manual_seed(0)
    preds = [torch.randn(2, 3, hw, hw, 85, generator=g) * 0.5 for hw in (80, 40, 20)]
    t0 = time.perf_counter()
    dec = [detect_ref.yolo_decode(p, detect_ref.anchors_wh(s)) for s, p in enumerate(preds)]
    detect_ref.post_process(dec, 0.25, 0.45, 100)
    dt = (time.perf_counter() - t0) / 2
    return {"value": 1.0 / dt, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "decode + two-stage NMS on 2 images (D18 worst case), oracle/detect_ref.py"}


def run_extras(args, dev, world, rank):
    """K2 microbenchmark, detection tail, and the hybrid_vision configurations (every rank runs them; sharded legs take
    the max over ranks)."""
    import torch
    import torch.distributed as dist
    from hvs_b200 import harness

    def rank_max(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tf_burst = float(peaks.get("bf16_tflops", 1590.0))
    out = {}
    k2 = harness.k2_microbench(dev)
    k2["frac_of_bf16_sustained"] = k2["tflops_fused"] / tf_sus
    k2["frac_of_bf16_burst"] = k2["tflops_fused"] / tf_burst
    k2["peak_tflops"] = {"sustained": tf_sus, "burst": tf_burst, "source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"}
    k2["speedup_vs_library_path"] = k2["ms_library_path"] / k2["ms_fused"]
    out["k2"] = k2
    worst = harness.detect_tail(dev, 64, 0.0)
    real = harness.detect_tail(dev, 64, -4.0)
    hbm = load_peaks()[0]
    for d in (worst, real):
        d["decode_frac_of_hbm"] = d["decode_GBps"] / hbm
    out["detect"] = {"worst_case_D18": worst, "objectness_bias_-4": real}
    torch.cuda.empty_cache()
    if args.skip_hybrid:
        return out
    import copy
    from hvs_b200.hybrid_vision import to_channels_last
    train_model = harness.build_model(dev, seed=0)
    model = copy.deepcopy(train_model)                     # inference copy: eval-mode BatchNorm folded into the convolutions
    folded = harness.fold_batchnorm_for_inference(model.eval())
    to_channels_last(model)
    harness.cast_weights_for_bf16_inference(model)         # the casts autocast would launch on every call, done once
    per = 64 // world
    inf = harness.inference_sharded(model, dev, world, rank, 64, 640)
    inf_real = harness.inference_sharded(model, dev, world, rank, 64, 640, objectness_bias=-4.0, steps=3, warmup=1)
    inf_e2e = harness.inference_sharded(model, dev, world, rank, 64, 640, steps=3, warmup=1, host_input=True)
    ms_inf, ms_real, ms_e2e = rank_max(inf["ms_per_step"]), rank_max(inf_real["ms_per_step"]), rank_max(inf_e2e["ms_per_step"])
    ips = 64 / (ms_inf * 1e-3)
    hv = {"inference": {"workload": "BASELINE configs[2]: batch 64 at 640x640, bf16 autocast + channels_last, decode + two-stage NMS (conf 0.25, iou 0.45, max 100), sharded 64/N, no collective (strong scaling)",
                        "img_per_s": ips, "ms_per_batch": ms_inf, "images_per_gpu": per,
                        "model_tflops": ips * harness.FWD_GFLOP_PER_IMAGE_640 / 1e3,
                        "frac_of_bf16_sustained_per_gpu": ips * harness.FWD_GFLOP_PER_IMAGE_640 / 1e3 / world / tf_sus,
                        "img_per_s_objectness_bias_-4": 64 / (ms_real * 1e-3), "mean_detections": inf["mean_detections"],
                        "hvs_launches_per_step": inf["hvs_launches_per_step"], "conv_bn_pairs_folded": folded,
                        "e2e": {"img_per_s": 64 / (ms_e2e * 1e-3), "ms_per_batch": ms_e2e, "h2d_bytes_per_step": inf_e2e["h2d_bytes_per_step"],
                                "d2h_bytes_per_step": inf_e2e["d2h_bytes_per_step"], "note": "per rank: pinned host images -> H2D -> forward -> decode -> NMS -> detections D2H, all timed"}}}
    if world > 1:
        # the same step with 64 images on EVERY GPU (weak scaling), next to configs[2]'s strong-scaling split of one batch of 64
        inf_w = harness.inference_sharded(model, dev, world, rank, 64 * world, 640, steps=3, warmup=1)
        ms_w = rank_max(inf_w["ms_per_step"])
        hv["inference"]["weak_scaling_64_per_gpu"] = {"img_per_s": 64 * world / (ms_w * 1e-3), "ms_per_batch": ms_w, "global_batch": 64 * world}
    torch.cuda.empty_cache()
    stream = harness.streaming_latency(model, dev, frames=args.stream_frames)
    hv["streaming"] = dict(stream, workload="BASELINE configs[4]: batch-1 640x640 frames, whole forward + decode + NMS in one CUDA graph")
    del model
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    model = train_model
    tr = harness.training_ddp(model, dev, world, rank, args.train_batch, 640, use_graph=not args.train_eager)
    ms_tr = rank_max(tr["ms_per_step"])
    tips = world * tr["batch_per_gpu"] / (ms_tr * 1e-3)
    hv["training"] = dict(tr, workload="BASELINE configs[3]: bf16 autocast training, synthetic COCO-shaped dense targets, YOLOLoss, AdamW, "
                                       "DistributedDataParallel gradient all-reduce over NCCL (weak scaling)",
                          ms_per_step=ms_tr, img_per_s=tips, model_tflops=tips * 3 * harness.FWD_GFLOP_PER_IMAGE_640 / 1e3,
                          frac_of_bf16_sustained_per_gpu=tips * 3 * harness.FWD_GFLOP_PER_IMAGE_640 / 1e3 / world / tf_sus)
    out["hybrid_vision"] = hv
    out["hybrid_vision_img_per_s"] = ips
    out["hybrid_vision_train_img_per_s"] = tips
    del model
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8192
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    tps, dt, threads = cpu_reference_run(sample, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mHC layer microbenchmark n=4 C=512 fwd+bwd, 20 Sinkhorn iters (BASELINE configs[1])",
                   "tokens_per_step": sample, "note": "bounded sample of the 2^20-token workload; CPU throughput is flat in T"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} tokens fwd+bwd, oracle/mhc_ref.py (torch fp32 CPU), mean of {steps} steps"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.skip_extras and not args.skip_hybrid:
        hb = cpu_hybrid_baseline()
        line["hybrid_vision"] = {"cpu_baseline": hb, "inference": {"img_per_s": hb["value"], "workload": hb["sample"]}}
        line["hybrid_vision_img_per_s"] = hb["value"]
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--tokens", type=int, default=1 << 20, help="tokens per GPU")
    ap.add_argument("--impl", default="hvs_b200", choices=["hvs_b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="no k2 / detect / hybrid_vision legs")
    ap.add_argument("--skip-hybrid", action="store_true", help="no hybrid_vision legs")
    ap.add_argument("--stream-frames", type=int, default=300)
    ap.add_argument("--train-batch", type=int, default=16)
    ap.add_argument("--train-eager", action="store_true", help="training step launched eagerly instead of one CUDA graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import hvs_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries ONE JSON line: everything libraries print (NCCL's version banner, ...) goes to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # the DDP training step is captured in a CUDA graph
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    T = args.tokens

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(T, N_STREAMS, CHANNELS, generator=g, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(T, N_STREAMS, CHANNELS, generator=g, device=dev, dtype=torch.bfloat16)
    gp = torch.Generator(device=dev).manual_seed(0)                      # identical parameters on every rank
    phi = torch.randn(N_STREAMS * CHANNELS, LOGITS, generator=gp, device=dev) * 0.02
    bias = torch.zeros(LOGITS, device=dev)
    alpha = torch.full((3,), 0.01, device=dev)
    scale = torch.ones(N_STREAMS * CHANNELS, device=dev)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    saved = hvs_b200.ops.new_saved(x)
    lib = hvs_b200.load_library()
    ws = torch.empty(int(lib.hvs_mhc_stream_bwd_saved_workspace(T, N_STREAMS, CHANNELS)), dtype=torch.uint8, device=dev)

    def step():                                                          # the training path of hvs_b200.StreamMHC
        hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved)
        grads = hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, out=dx, workspace=ws)
        if world > 1:                                                    # DDP-style parameter-gradient exchange
            flat = torch.cat([grads[k].reshape(-1) for k in ("dphi", "dbias", "dalpha", "dscale")])
            dist.all_reduce(flat)
        return grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    launches0 = hvs_b200._lib.launch_count()
    lib.hvs_mhc_stream_profile(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    barrier()
    t_wall = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        step()
    ev1.record()
    barrier()
    buf = (ctypes.c_float * 4)()
    lib.hvs_mhc_stream_kernel_ms(buf)                    # mean per-kernel event durations over the timed steps (read after them:
    kms.append(list(buf))                                # the host does not wait inside the timed region)
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    lib.hvs_mhc_stream_profile(0)
    launches = hvs_b200._lib.launch_count() - launches0
    step_ms = ev0.elapsed_time(ev1) / args.steps         # exactly K steps, including the gaps between them
    t = torch.tensor([step_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = world * T / (step_ms * 1e-3)

    # ---- the same step with HVS_MHC_ADAPTIVE_ITERS (opt-in: the Sinkhorn loops stop at convergence to 2^-20; forward results
    #      within 2e-6, gradients within 1e-5).  Reported NEXT to the headline, which runs all 20 iterations.
    def step_adaptive():
        hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, out=y, saved=saved, adaptive=True)
        hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, out=dx, workspace=ws, adaptive=True)
    time.sleep(1.5)                                      # same starting clocks as the headline loop (back-to-back loops drift with the power cap)
    for _ in range(warmup):
        step_adaptive()
    barrier()
    lib.hvs_mhc_stream_profile(1)
    ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea0.record()
    for _ in range(args.steps):
        step_adaptive()
    ea1.record()
    barrier()
    abuf = (ctypes.c_float * 4)()
    lib.hvs_mhc_stream_kernel_ms(abuf)
    lib.hvs_mhc_stream_profile(0)
    ta = torch.tensor([ea0.elapsed_time(ea1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    adaptive_ms = float(ta.item())
    adaptive_kernels = list(abuf)

    # ---- end to end through the host-buffer entry (pinned host memory, copies inside the timed region)
    layer = hvs_b200.StreamMHC(device=dev)
    with torch.no_grad():
        layer.phi.copy_(phi); layer.bias.copy_(bias); layer.alpha.copy_(alpha); layer.rms_scale.copy_(scale)
    del y, dx, ws
    torch.cuda.empty_cache()
    xh = x.cpu().pin_memory(); dyh = dy.cpu().pin_memory()
    yh = torch.empty_like(xh).pin_memory(); dxh = torch.empty_like(xh).pin_memory()
    hvs_b200.stream_mhc_fwd_bwd_host(xh[: 1 << 16], dyh[: 1 << 16], layer, yh[: 1 << 16], dxh[: 1 << 16])   # warm-up
    barrier()
    e2e_ts = []
    for _ in range(args.e2e_steps):
        t0 = time.perf_counter()
        gh = hvs_b200.stream_mhc_fwd_bwd_host(xh, dyh, layer, yh, dxh)
        torch.cuda.synchronize()
        e2e_ts.append(time.perf_counter() - t0)
    e2e_dt = sum(e2e_ts) / len(e2e_ts) if e2e_ts else float('nan')
    te = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * T / float(te.item())
    row_bytes = N_STREAMS * CHANNELS * 2
    grad_bytes = 4 * (N_STREAMS * CHANNELS * LOGITS + LOGITS + 3 + N_STREAMS * CHANNELS)
    del x, dy, xh, dyh, yh, dxh, saved
    torch.cuda.empty_cache()
    extras = {}
    if not args.skip_extras:
        extras = run_extras(args, dev, world, rank)

    if rank == 0:
        peak, peak_src = load_peaks()
        fwd_ms = sum(k[0] for k in kms) / len(kms)
        bwd_ms = sum(k[1] for k in kms) / len(kms)
        fin_ms = sum(k[3] for k in kms) / len(kms)
        achieved = BWD_BYTES * T / (bwd_ms * 1e-3) / 1e9
        traffic = load_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 in/out, fp32 accumulate and Sinkhorn", "data": "synthetic",
            "config": {"workload": "mHC layer microbenchmark n=4 C=512 bf16 fwd+bwd, 20 Sinkhorn iters (BASELINE configs[1])",
                       "tokens_per_gpu": T, "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": "inputs (4.3 GB each) are larger than L2; no flush needed",
                       "params": "phi~N(0,0.02^2), bias=0, alpha=0.01, rms_scale=1",
                       "path": "training path: forward saves 112 B/token of statistics, backward = one fused kernel (dx + all parameter gradients) + a 2048-row finalize"},
            "hbm_gbs": {"fwd_bwd_algorithmic": (FWD_BYTES + BWD_BYTES) * T * world / (step_ms * 1e-3) / 1e9,
                        "frac_of_peak": (FWD_BYTES + BWD_BYTES) * T / (step_ms * 1e-3) / 1e9 / peak, "peak": peak,
                        "peak_source": peak_src, "frac_of_nominal_8000": (FWD_BYTES + BWD_BYTES) * T / (step_ms * 1e-3) / 1e9 / 8000.0},
            "kernels_ms": {"mhc_stream_fwd_kernel": fwd_ms, "mhc_stream_bwd_fused_kernel": bwd_ms,
                           "mhc_stream_bwd_finalize_kernel": fin_ms,
                           "fwd_GBs": FWD_BYTES * T / (fwd_ms * 1e-3) / 1e9, "fwd_frac": FWD_BYTES * T / (fwd_ms * 1e-3) / 1e9 / peak},
            "roofline": {"kernel": "mhc_stream_bwd_fused_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback (B200_PROFILING.md)",
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get("mhc_stream_bwd_fused_kernel_bytes_per_launch"),
                         "traffic_note": (traffic or {}).get("note")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * T * row_bytes,
                    "d2h_bytes_per_step": 2 * T * row_bytes + grad_bytes, "ms_per_step": e2e_dt * 1e3,
                    "api": "hvs_b200.stream_mhc_fwd_bwd_host (pinned host buffers, 3-stream chunk pipeline)"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
            "adaptive_sinkhorn": {
                "note": "OPT-IN variant, not the headline: HVS_MHC_ADAPTIVE_ITERS stops each warp's Sinkhorn loop at the first iteration that changes nothing "
                        "by more than 2^-20 (at this config's logit scale after 3 of the 20 iterations); coefficients within 2e-6 of the full run (bound: 1e-5), "
                        "gradients within 1e-5 (tests/test_gpu_mhc_stream.py::test_adaptive_iterations_*).  The headline `value` runs all 20 iterations like the reference",
                "ms_per_step": adaptive_ms, "value": world * T / (adaptive_ms * 1e-3), "unit": UNIT,
                "kernels_ms": {"mhc_stream_fwd_kernel": adaptive_kernels[0], "mhc_stream_bwd_fused_kernel": adaptive_kernels[1]},
                "bwd_frac_of_peak": BWD_BYTES * T / (adaptive_kernels[1] * 1e-3) / 1e9 / peak if adaptive_kernels[1] > 0 else None,
                "fwd_bwd_frac_of_peak": (FWD_BYTES + BWD_BYTES) * T / (adaptive_ms * 1e-3) / 1e9 / peak,
                "fwd_bwd_frac_of_nominal_8000": (FWD_BYTES + BWD_BYTES) * T / (adaptive_ms * 1e-3) / 1e9 / 8000.0},
        }
        if world == 1 and not args.no_cpu_baseline:
            tps, dt, threads = cpu_reference_run(8192, 2, 1)
            line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "8192 tokens fwd+bwd, oracle/mhc_ref.py (torch fp32 CPU), mean of 2 steps after 1 warm-up"}
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline and "hybrid_vision" in line:
            line["hybrid_vision"]["cpu_baseline"] = cpu_hybrid_baseline()
            if "detect" in line:
                line["detect"]["cpu_baseline"] = cpu_detect_baseline()
        if world == 1 and not args.no_cpu_baseline and "k2" in line:
            line["k2"]["cpu_baseline"] = cpu_k2_baseline()
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
