#!/usr/bin/env python
"""Benchmark of the FS-EEND hot path (encoder + attractor-decoder forward -> per-frame logits).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1] shape): B=64 sequences x T=500 frames x D=345, S=6 attractor slots
(4 speakers + silence + no-speaker), synthetic N(0,1) features (seed 777), default-init weights (seed 0).
A "step" is one forward over one batch.  Metric: audio frames/s = N_gpus * B * T * steps / time, timed with
CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks.  Multi-GPU: the path
shards by sequence with no data-path collective (weak scaling, one process per GPU).

One JSON line on stdout (rank 0).  See DESIGN.md §measurement for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "fs-eend_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

B, T, DIN, S, D, H = 64, 500, 345, 6, 256, 4
ENC_L, DEC_L, FF = 4, 2, 2048
METRIC = "audio frames/sec (enc+attractor fwd) at T=500, 4-spk"
UNIT = "frames/s"


# ------------------------------------------------------------------------------------------ flop model
def algorithmic_flops(name: str, Bn: int, Tn: int, Sn: int) -> float:
    """ALGORITHMIC FLOPs of one launch of the named kernel (1 MAC = 2 FLOP, causal-exact
    attention).  Names are the profiling tags of fs_model.cu."""
    Me, Md = Bn * Tn, Bn * Tn * Sn
    attn_seq_head = 4.0 * 64 * Tn * (Tn + 1) / 2          # QK^T + PV, causal-exact
    table = {
        "enc.gemm_in_ln": 2.0 * Me * DIN * D,
        "enc.gemm_qkv": 2.0 * Me * D * 3 * D,
        "enc.attn_causal": attn_seq_head * H * Bn,
        "enc.gemm_out_ln": 2.0 * Me * D * D,
        "enc.gemm_ffn1": 2.0 * Me * D * FF,
        "enc.gemm_ffn2_ln": 2.0 * Me * D * FF,
        "gemm_conv_l2": 2.0 * Me * D * D * 19,
        "gemm_convert": 2.0 * Me * D * D,
        "dec.gemm_qkv1": 2.0 * Md * D * 3 * D,
        "dec.attn_causal": attn_seq_head * H * Bn * Sn,
        "dec.gemm_out1_ln": 2.0 * Md * D * D,
        "dec.gemm_qkv2": 2.0 * Md * D * 3 * D,
        "dec.spk_attn": 4.0 * 64 * H * Sn * Sn * Me,
        "dec.spk_fused": 2.0 * Md * D * 3 * D + 4.0 * 64 * H * Sn * Sn * Me,
        "dec.gemm_out2_ln": 2.0 * Md * D * D,
        "dec.gemm_ffn1": 2.0 * Md * D * FF,
        "dec.gemm_ffn2_ln": 2.0 * Md * D * FF,
        "dec.ffn_fused": 4.0 * Md * D * FF,
        "enc.ffn_fused": 4.0 * Me * D * FF,
    }
    return table.get(name, 0.0)


def total_flops(Bn, Tn, Sn):
    per_layer_enc = sum(algorithmic_flops(n, Bn, Tn, Sn) for n in
                        ("enc.gemm_qkv", "enc.attn_causal", "enc.gemm_out_ln", "enc.gemm_ffn1", "enc.gemm_ffn2_ln"))
    per_layer_dec = sum(algorithmic_flops(n, Bn, Tn, Sn) for n in
                        ("dec.gemm_qkv1", "dec.attn_causal", "dec.gemm_out1_ln", "dec.gemm_qkv2", "dec.spk_attn",
                         "dec.gemm_out2_ln", "dec.gemm_ffn1", "dec.gemm_ffn2_ln"))
    return (algorithmic_flops("enc.gemm_in_ln", Bn, Tn, Sn) + ENC_L * per_layer_enc +
            algorithmic_flops("gemm_conv_l2", Bn, Tn, Sn) + algorithmic_flops("gemm_convert", Bn, Tn, Sn) +
            DEC_L * per_layer_dec + 2.0 * Bn * Tn * Sn * D)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                pk = json.load(f)
            return {"hbm_gbs": float(pk.get("hbm_gbs", 6650.0)),
                    "tflops": float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))),
                    "tflops_burst": float(pk.get("bf16_tflops", 1590.0)), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_burst": 1590.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms (a timed
    region of a few hundred ms is over before an `nvidia-smi -lms` child process has even started); falls back to
    `nvidia-smi` polling when the NVML binding is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons = index, [], 0, set()
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def _nvml_loop(self, nv, h):
        masks = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                 else nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                         getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        while not self._stop.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for n, m in masks.items():
                    if r & m:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_loop(self):
        fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._physical_index()), f"--query-gpu={fields}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                self.sm.append(int(p[0]))
                self.mx = max(self.mx, int(p[1]))
                for n, v in zip(self.NAMES, p[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                time.sleep(0.05)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.source = "nvidia-smi"
            self._thread = threading.Thread(target=self._smi_loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------ CPU arm
def pick_cpu_threads(sd, src, lens, cfg):
    """The reference's CPU path is torch intra-op parallel; on a many-core host the best thread count for these
    small matrices is well below the core count.  Probe a few counts on one sequence and keep the fastest."""
    import torch
    from oracle import fs_eend_oracle as O
    O.USE_SDPA = True   # the fused CPU attention path the reference's nn.MultiheadAttention takes
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for th in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(th)
        with torch.no_grad():
            O.test(sd, src[:1], lens[:1], S, cfg)
            t0 = time.perf_counter()
            O.test(sd, src[:1], lens[:1], S, cfg)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = th, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_throughput(seconds_budget: float, n_seq: int, threads: int = 0):
    """Times the CPU restatement of the reference path (oracle port, same ATen CPU kernels the reference's
    torch modules dispatch to) on `n_seq`-sequence samples of the workload.  Returns (frames/s, s, reps, threads)."""
    import torch
    from oracle import fs_eend_oracle as O
    sd = O.random_state_dict(seed=0, trained_like=False)
    src, lens = O.synthetic_features(n_seq, T)
    cfg = O.Cfg()
    threads = threads or pick_cpu_threads(sd, src, lens, cfg)
    torch.set_num_threads(threads)
    with torch.no_grad():
        O.test(sd, src[:1], lens[:1], S, cfg)       # warm-up
        t0 = time.perf_counter()
        reps = 0
        while True:
            O.test(sd, src, lens, S, cfg)
            reps += 1
            el = time.perf_counter() - t0
            if el > seconds_budget or reps >= 50:
                break
    return n_seq * T * reps / el, el, reps, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n_seq = 4
    # warm-up steps then `steps` timed steps, each a bounded n_seq-sequence sample
    from oracle import fs_eend_oracle as O
    sd = O.random_state_dict(seed=0, trained_like=False)
    src, lens = O.synthetic_features(n_seq, T)
    cfg = O.Cfg()
    threads = pick_cpu_threads(sd, src, lens, cfg)
    steps = min(args.steps, 20)
    with torch.no_grad():
        for _ in range(min(args.warmup, 2)):
            O.test(sd, src, lens, S, cfg)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.test(sd, src, lens, S, cfg)
        el = time.perf_counter() - t0
    value = n_seq * T * steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"FS-EEND fwd, {n_seq}-sequence sample of B={B} T={T} D={DIN} S={S} per step (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x {n_seq} sequences x {T} frames, torch CPU fp32, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ secondary workloads
def secondary_measurements(dev):
    """The other BASELINE.json configurations, measured in the same run (rank 0, N = 1) with their own CPU figure beside
    them: LS-EEND batch inference at B=16 x T=2000 x S=10 (configs[2]) in both precision modes (device-resident, end to
    end through fseend_ls_forward_host, logit error against the CPU oracle on one recording), frame-by-frame latency
    of FS-EEND and LS-EEND (configs[4]: the one-hour recording is per-frame latency x 36000, the LS step is O(1))."""
    import torch
    from oracle import fs_eend_oracle as FO
    from oracle import ls_eend_oracle as LO
    from nnet.model.onl_conformer_retention_enc_1dcnn_tfm_retention_enc_linear_non_autoreg_pos_enc_l2norm_emb_loss_mask import (
        OnlineConformerRetentionDADiarization)
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization
    from nnet.model.streaming_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import StreamingTransformerEDADiarization
    from nnet.utils.copy_params import copy_params_from_masked_to_streaming

    def ev_time(fn, warm, iters):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    ncpu = os.cpu_count() or 1
    cpu_threads = min(ncpu, 16)
    torch.set_num_threads(cpu_threads)
    # ---------------- LS-EEND batch, BASELINE configs[2]
    LB, LT, LS_ = 16, 2000, 10
    sd = LO.random_state_dict(seed=4, trained_like=True)
    ls = OnlineConformerRetentionDADiarization(
        n_speakers=8, in_size=DIN, n_units=D, n_heads=H, enc_n_layers=4, dec_n_layers=2, dropout=0.1, max_seqlen=1000,
        recurrent_chunk_size=500, feed_forward_expansion_factor=4, dec_dim_feedforward=FF, conv_kernel_size=16)
    ls.load_state_dict(sd, strict=True)
    ls = ls.to(dev).eval()
    src, lens = FO.synthetic_features(LB, LT)
    x_host = torch.cat(src).contiguous().pin_memory()
    x_dev = x_host.to(dev)
    with torch.no_grad():
        t0 = time.perf_counter()
        ref0 = LO.test(sd, src[:1], lens[:1], LS_, LO.Cfg())[0][0]
        cpu_s = time.perf_counter() - t0
    rec = {"workload": f"LS-EEND fwd B={LB} T={LT} S={LS_} (8-spk), retention chunk 500",
           "cpu_baseline": {"value": LT / cpu_s, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                            "sample": f"1 recording x {LT} frames, oracle port, torch CPU fp32, {cpu_threads} threads"}}
    nat = ls.native()
    out_host = torch.empty(LB, LT, LS_).pin_memory()
    for mode in ("fp32", "fp16"):
        nat.set_precision(mode)
        y = nat.forward(x_dev, lens, LS_)[0]
        err = (y[0].cpu() - ref0).abs()
        ms = ev_time(lambda: nat.forward(x_dev, lens, LS_), 2, 5)
        nat.forward_host(x_host, lens, LS_, out=out_host)
        t0 = time.perf_counter()
        for _ in range(3):
            nat.forward_host(x_host, lens, LS_, out=out_host)
        e2e_s = (time.perf_counter() - t0) / 3
        rec[mode] = {"dtype": "fp32 activations, split fp16 hi+lo tensor-core operands (3 MMAs)" if mode == "fp32"
                     else "fp16 operands and activations, fp32 accumulate",
                     "ms_per_forward": ms, "frames_per_s": LB * LT / ms * 1e3,
                     "e2e_frames_per_s": LB * LT / e2e_s, "e2e_api": "fseend_ls_forward_host",
                     "h2d_bytes": LB * LT * DIN * 4, "d2h_bytes": LB * LT * LS_ * 4,
                     "launches": nat.launches_per_forward,
                     "logit_err_vs_oracle": {"max": float(err.max()), "median": float(err.median()),
                                             "p99": float(err.flatten().kthvalue(int(0.99 * err.numel())).values)}}
    out["ls_batch_B16_T2000_S10"] = rec
    # ---------------- LS-EEND frame by frame, B = 1, S = 10 (BASELINE configs[4]: T = 36000 = this latency x 36000)
    n_cpu_frames = 40
    with torch.no_grad():
        t0 = time.perf_counter()
        LO.stream_all(sd, src[0][None, :n_cpu_frames], LS_, LO.Cfg())
        cpu_frame = (time.perf_counter() - t0) / (n_cpu_frames + 9)
    rec = {"workload": "LS-EEND one-step (recurrent) inference, B=1, S=10; 1 hour = 36000 frames of 100 ms",
           "cpu_baseline": {"ms_per_frame": cpu_frame * 1e3, "real_time_factor": cpu_frame / 0.1, "cores": cpu_threads,
                            "kind": "port", "sample": f"{n_cpu_frames} frames + flush, oracle port"}}
    xt = torch.randn(1, DIN, device=dev)
    for mode in ("fp32", "fp16"):
        nat.set_precision(mode)
        st = ls.new_stream(1, LS_)
        for _ in range(40):
            st.step(xt)
        torch.cuda.synchronize()
        n = 400
        t0 = time.perf_counter()
        for _ in range(n):
            st.step(xt)
        torch.cuda.synchronize()
        lat = (time.perf_counter() - t0) / n
        rec[mode] = {"ms_per_frame": lat * 1e3, "real_time_factor": lat / 0.1, "one_hour_T36000_seconds": lat * 36009,
                     "frames_timed": n}
        del st
    out["ls_stream_B1_S10"] = rec
    nat.set_precision("fp32")
    # ---------------- FS-EEND frame by frame, B = 1, S = 6
    fsd = FO.random_state_dict(seed=0, trained_like=False)
    kw = dict(in_size=DIN, n_units=D, n_heads=H, enc_n_layers=ENC_L, dec_n_layers=DEC_L, dropout=0.1, has_mask=True,
              max_seqlen=T, dec_dim_feedforward=FF)
    fs = OnlineTransformerDADiarization(n_speakers=4, **kw)
    fs.load_state_dict(fsd, strict=True)
    fs = fs.to(dev).eval()
    sfs = StreamingTransformerEDADiarization(**kw).to(dev).eval()
    copy_params_from_masked_to_streaming(fs, sfs)
    xt3 = torch.randn(1, 1, DIN, device=dev)
    for _ in range(40):                 # first use of every kernel (lazy module loading), graph capture
        sfs.test(xt3, S)
    sfs.reset()                         # new recording
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        sfs.test(xt3, S)
    torch.cuda.synchronize()
    fs_lat = (time.perf_counter() - t0) / 500
    from fseend_b200.native import FsStream
    fst = FsStream(fs.native(), 1, S)
    xt2 = torch.randn(1, DIN, device=dev)
    for _ in range(30):
        fst.step(xt2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(470):
        fst.step(xt2)
    torch.cuda.synchronize()
    fs_native_lat = (time.perf_counter() - t0) / 470
    fsrc, _ = FO.synthetic_features(1, 60)
    with torch.no_grad():
        t0 = time.perf_counter()
        FO.stream_all(fsd, fsrc[0][None], S, FO.Cfg())
        fs_cpu = (time.perf_counter() - t0) / 69
    out["fs_stream_B1_S6"] = {"workload": "FS-EEND frame-by-frame, B=1, S=6, first 500 frames (attention over the growing cache)",
                              "ms_per_frame": fs_lat * 1e3, "real_time_factor": fs_lat / 0.1,
                              "api": "StreamingTransformerEDADiarization.test (the reference's frame-loop call)",
                              "native_step_ms_per_frame": fs_native_lat * 1e3,
                              "cpu_baseline": {"ms_per_frame": fs_cpu * 1e3, "cores": cpu_threads, "kind": "port",
                                               "sample": "60 frames + flush, oracle port (first 60 frames: shorter cache "
                                                         "than the GPU figure's 500)"}}
    # ---------------- FS-EEND training step (started): native forward/backward kernels vs torch eager
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_step_bench
        rec = train_step_bench.measure(steps=3, warmup=2)
        # CPU figure beside it: the oracle restatement's forward + both losses + torch-autograd backward on the host cores
        # (float32, 4 chunks x 500 frames, 6 label classes; BatchNorm on running statistics; no optimizer step)
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and not k.endswith(".pe"))
              for k, v in FO.random_state_dict(seed=0).items()}
        csrc, clens = FO.synthetic_features(4, T)
        ctgt = FO.synthetic_labels(3, clens, [6] * 4)
        t0 = time.perf_counter()
        o, el, _, _ = FO.forward(sd, csrc, ctgt, clens, FO.Cfg())
        bce = sum(torch.nn.functional.binary_cross_entropy_with_logits(y, t) * len(y) for y, t in zip(o, ctgt)) / sum(clens)
        (bce + el).backward()
        cpu_s = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 4 * T / cpu_s, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                               "sample": f"4 chunks x {T} frames, oracle forward + torch autograd backward, fp32, {cpu_threads} threads, {cpu_s:.1f} s"}
        out["fs_train_step_B64_T500"] = rec
    except Exception as e:  # the headline must not depend on the secondary training measurement
        out["fs_train_step_B64_T500"] = {"error": repr(e)[:300]}
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    torch.manual_seed(0)
    model = OnlineTransformerDADiarization(
        n_speakers=4, in_size=DIN, n_units=D, n_heads=H, enc_n_layers=ENC_L, dec_n_layers=DEC_L, dropout=0.1,
        has_mask=True, max_seqlen=T, dec_dim_feedforward=FF).cuda().eval()
    native = model.native()
    for kv in filter(None, os.environ.get("FSEEND_OPTS", "").split(",")):   # e.g. FSEEND_OPTS=ffn=2,spk=0 (kernel variants)
        k, v = kv.split("=")
        native.set_option(k, int(v))
    lens = [T] * B
    gen = torch.Generator(device="cpu").manual_seed(777 + rank)
    n_buf = 4   # 4 x 44 MB of inputs > 126 MB L2: inputs are never L2-warm
    xs_host = [torch.randn(B * T, DIN, generator=gen).pin_memory() for _ in range(n_buf)]
    xs = [x.to(dev) for x in xs_host]
    out_host = torch.empty(B, T, S).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for i in range(args.warmup):
        native.forward(xs[i % n_buf], lens, S)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        native.forward(xs[i % n_buf], lens, S)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = native.launches_per_forward * args.steps * world      # whole job: every rank launches the same sequence

    # ---- sustained figure: the same loop for >= 200 steps (a 20-step region is ~45 ms: boost clocks, no power cap yet)
    sustained = None
    if rank == 0 and world == 1 and args.steps < 200 and not args.no_secondary:
        n_sus = 300
        s_sampler = ClockSampler(local)
        s_sampler.start()
        sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        sv0.record()
        for i in range(n_sus):
            native.forward(xs[i % n_buf], lens, S)
        sv1.record()
        torch.cuda.synchronize()
        s_ms = sv0.elapsed_time(sv1)
        sustained = {"steps": n_sus, "value": B * T * n_sus / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / n_sus,
                     "clocks": s_sampler.stop()}

    # ---- end-to-end through the host-buffer C-ABI entry points (H2D + forward + D2H every step, pinned host buffers)
    # (a) pipelined: fseend_fs_forward_host_async / fseend_fs_host_wait, two calls in flight — the copy of step i+1
    #     overlaps the kernels of step i; every step's input still crosses PCIe and every step's logits are read back
    #     inside the timed region.  (b) blocking: one fseend_fs_forward_host call per step (copy + compute serialised
    #     except for the two-chunk overlap inside the call).
    e2e_steps = max(3, args.steps)
    out_hosts = [torch.empty(B, T, S).pin_memory() for _ in range(2)]
    for i in range(2):
        native.host_wait(native.forward_host_async(xs_host[i % n_buf], lens, S, out_hosts[i % 2]))
    barrier()
    t0 = time.perf_counter()
    prev = None
    for i in range(e2e_steps):
        tk = native.forward_host_async(xs_host[i % n_buf], lens, S, out_hosts[i % 2])
        if prev is not None:
            native.host_wait(prev)
        prev = tk
    native.host_wait(prev)
    e2e_s = time.perf_counter() - t0
    sync_steps = max(3, min(args.steps, 20))
    for i in range(2):
        native.forward_host(xs_host[i % n_buf], lens, S, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(sync_steps):
        native.forward_host(xs_host[i % n_buf], lens, S, out=out_host)
    torch.cuda.synchronize()
    e2e_sync_s = (time.perf_counter() - t0) / sync_steps

    from fseend_b200.parallel import max_over_ranks
    ms, e2e_ms, e2e_sync_ms = max_over_ranks([ms, e2e_s * 1e3, e2e_sync_s * 1e3], dev)

    # ---- per-kernel roofline (rank 0, separate profiled passes: CUDA events around every launch)
    roof, roof_attn, prof_table = None, None, None
    if rank == 0:
        native.set_profiling(True)
        n_prof = 3
        for i in range(n_prof):
            native.forward(xs[i % n_buf], lens, S)
        torch.cuda.synchronize()
        prof = native.get_profile()
        native.set_profiling(False)
        pk = measured_peaks()
        prof_table = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] // n_prof,
                          "tflops": (algorithmic_flops(k, B, T, S) / (v[0] / v[1] * 1e-3) / 1e12)
                          if algorithmic_flops(k, B, T, S) else None} for k, v in prof.items()}
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass
        dom = max(prof, key=lambda k: prof[k][0])
        d_ms = prof[dom][0] / prof[dom][1]
        fl = algorithmic_flops(dom, B, T, S)
        ach = fl / (d_ms * 1e-3) / 1e12
        # The kernel is timed alone (events around single launches in a 3-forward pass: boost clocks, no power cap), so
        # the BURST cuBLAS figure is the denominator that applies; the fraction of the sustained figure is given beside it.
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["tflops_burst"], "peak_kind": "burst (kernel timed in isolation)",
                "frac_of_sustained_peak": ach / pk["tflops"], "peak_sustained": pk["tflops"],
                "traffic": traffic.get(dom), "peak_source": pk["source"],
                "avg_launch_ms": d_ms, "algorithmic_flops_per_launch": fl}
        for k in ("dec.attn_causal",):
            if k in prof:
                a_ms = prof[k][0] / prof[k][1]
                a = algorithmic_flops(k, B, T, S) / (a_ms * 1e-3) / 1e12
                roof_attn = {"kernel": k, "bound": "tensor", "achieved": a, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                             "frac": a / pk["tflops_burst"], "peak_kind": "burst (kernel timed in isolation)",
                             "frac_of_sustained_peak": a / pk["tflops"], "avg_launch_ms": a_ms, "flops": "causal-exact",
                             "traffic": traffic.get(k)}

    # ---- N > 1: the one collective of the reference's training step (DDP gradient all-reduce, train_dia.py:145-160),
    # measured on its own: 9,974,450 fp32 gradients = 39.9 MB over NCCL.  NOT part of `value` (the forward path has no
    # exchange step; the backward kernels that would produce these gradients are not built) — it states
    # what the collective costs next to a forward of `ms_per_step`.
    grad_allreduce = None
    if world > 1:
        gbuf = torch.zeros(9_974_450, device=dev, dtype=torch.float32)
        for _ in range(3):
            dist.all_reduce(gbuf)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        a0.record()
        n_ar = 20
        for _ in range(n_ar):
            dist.all_reduce(gbuf)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = max_over_ranks([a0.elapsed_time(a1) / n_ar], dev)[0]
        nbytes = gbuf.numel() * 4
        grad_allreduce = {"bytes": nbytes, "ms": ar_ms, "algbw_gbs": nbytes / (ar_ms * 1e-3) / 1e9,
                          "busbw_gbs": nbytes / (ar_ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
                          "note": "NCCL all-reduce of the FS-EEND gradient volume alone; not included in value"}
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, el, reps, threads = cpu_reference_throughput(12.0, 4)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"{reps} x 4 sequences x {T} frames in {el:.1f} s (oracle port, torch CPU fp32, "
                         f"best of 8/16/32/64/all threads)"}

    secondary = None
    if world == 1 and not args.no_secondary:
        try:
            secondary = secondary_measurements(dev)
        except Exception as e:      # the headline line must not be lost to a secondary workload
            secondary = {"error": f"{type(e).__name__}: {e}"}

    frames = world * B * T
    value = frames * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16", "numerics": "fp16 tensor-core operands (same rate as bf16), fp32 accumulate / softmax / LayerNorm; "
                                     "logits within 2.5e-4 of the fp32 reference (bf16 operands measured 1.1-1.6e-3)",
        "data": "synthetic",
        "config": {"workload": f"FS-EEND enc+attractor fwd B={B}/GPU T={T} D={DIN} S={S} (4-spk), batch sharded by sequence",
                   "l2": "inputs rotate over 4 x 44 MB buffers (> 126 MB L2); per-step activations 1.6 GB >> L2",
                   "parallelism": f"dp{world} (no data-path collective)"},
        "tflops_algorithmic": total_flops(B, T, S) * world * args.steps / (ms * 1e-3) / 1e12,
        "e2e": {"value": frames * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": world * B * T * DIN * 4, "d2h_bytes_per_step": world * B * T * S * 4,
                "steps": e2e_steps, "api": "fseend_fs_forward_host_async + fseend_fs_host_wait (pinned host buffers, two "
                                             "calls in flight: H2D of step i+1 overlaps the kernels of step i)",
                "blocking_call": {"value": frames / (e2e_sync_ms * 1e-3), "unit": UNIT, "steps": sync_steps,
                                  "api": "fseend_fs_forward_host (one blocking call per step)"}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_attention": roof_attn,
        "cpu_baseline": cpu,
        "sustained": sustained,
        "secondary": secondary,
        "grad_allreduce": grad_allreduce,
        "kernels": prof_table,
        "options": os.environ.get("FSEEND_OPTS", "default"),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ training arm (--train)
def run_train(args):
    """one FS-EEND training step — forward, standard_loss + emb loss, backward through this library's
    kernels, gradient all-reduce (torch DistributedDataParallel over NCCL, 39.9 MB of fp32 gradients), Adam — at the
    BASELINE config-1 shape per GPU (weak scaling).  Prints ONE JSON line; `allreduce.exposed_ms` is the step time with the
    collective minus the step time under `no_sync()` (what the all-reduce adds after overlap with the backward)."""
    import contextlib
    import torch
    import torch.distributed as dist
    from fseend_b200.loss import standard_loss
    from nnet.model.onl_tfm_enc_1dcnn_enc_linear_non_autoreg_pos_enc_l2norm import OnlineTransformerDADiarization

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    model = OnlineTransformerDADiarization(
        n_speakers=4, in_size=DIN, n_units=D, n_heads=H, enc_n_layers=ENC_L, dec_n_layers=DEC_L, dropout=0.1,
        has_mask=True, max_seqlen=T, dec_dim_feedforward=FF).cuda().train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)
    g = torch.Generator(device="cuda").manual_seed(777 + rank)
    n_spk = 4
    src = [torch.randn(T, DIN, device="cuda", generator=g) for _ in range(B)]
    act = [(torch.rand(T, n_spk, device="cuda", generator=g) < 0.3).float() for _ in range(B)]
    lab = [torch.cat([1 - t.max(-1, keepdim=True)[0], t, torch.zeros(T, 1, device="cuda")], -1) for t in act]   # + silence, + none
    lens = [T] * B
    n_grad = sum(p.numel() for p in model.parameters() if p.requires_grad)

    def step(sync=True):
        ctx = contextlib.nullcontext() if (sync or world == 1) else net.no_sync()
        with ctx:
            opt.zero_grad(set_to_none=True)
            out, emb_loss, _, _ = net(src, lab, lens)
            loss = standard_loss(out, lab) + emb_loss
            loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, sync):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step(sync)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), loss

    steps = min(args.steps, 50)
    for _ in range(max(args.warmup, 3)):
        step(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, loss = timed(steps, True)
    clocks = sampler.stop() if rank == 0 else None
    ms_nosync = None
    if world > 1:
        for _ in range(2):
            step(False)
        ms_nosync, _ = timed(steps, False)
    if rank == 0:
        frames = world * B * T
        line = {"metric": "training_frames_per_second", "value": frames * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (split fp16 hi+lo tensor-core operands, fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": f"FS-EEND training step, {B} chunks x {T} frames per GPU, {n_spk} speakers + 2 label columns, "
                                       "4 enc + 2 dec layers, dropout 0.1, Adam; inputs resident in HBM",
                           "parallelism": f"dp{world}" if world > 1 else "single"},
                "status": "SURVEY 8f N1 started: GEMM / attention / LayerNorm / BatchNorm / L2 / head forward+backward native; "
                          "emb loss, layout copies, residual dropout, optimizer and the DDP bucketing are torch's",
                "allreduce": None if world == 1 else {
                    "bytes": 4 * n_grad, "collective": "NCCL all-reduce issued by torch DDP (25 MB buckets, overlapped with backward)",
                    "ms_per_step_without_sync": ms_nosync / steps, "exposed_ms": (ms - ms_nosync) / steps},
                "loss": float(loss.detach()), "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the LS-EEND / streaming secondary workloads")
    ap.add_argument("--train", action="store_true", help="measure the training step (forward + backward + all-reduce + Adam) instead")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.train:
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
