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
    launches = native.launches_per_forward * args.steps

    # ---- end-to-end through the host-buffer C-ABI entry point (H2D + forward + D2H every step)
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(2):
        native.forward_host(xs_host[i % n_buf], lens, S, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        native.forward_host(xs_host[i % n_buf], lens, S, out=out_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    from fseend_b200.parallel import max_over_ranks
    ms, e2e_ms = max_over_ranks([ms, e2e_s * 1e3], dev)

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
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["tflops"], "traffic": traffic.get(dom), "peak_source": pk["source"],
                "avg_launch_ms": d_ms, "algorithmic_flops_per_launch": fl}
        for k in ("dec.attn_causal",):
            if k in prof:
                a_ms = prof[k][0] / prof[k][1]
                a = algorithmic_flops(k, B, T, S) / (a_ms * 1e-3) / 1e12
                roof_attn = {"kernel": k, "bound": "tensor", "achieved": a, "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": a / pk["tflops"], "avg_launch_ms": a_ms, "flops": "causal-exact"}

    if world > 1:
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
                "h2d_bytes_per_step": B * T * DIN * 4, "d2h_bytes_per_step": B * T * S * 4,
                "steps": e2e_steps, "api": "fseend_fs_forward_host (pinned host buffers)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_attention": roof_attn,
        "cpu_baseline": cpu,
        "kernels": prof_table,
        "options": os.environ.get("FSEEND_OPTS", "default"),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
