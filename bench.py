#!/usr/bin/env python
"""Headline benchmark: audio-seconds per second of the separator forward on N B200s.

    python bench.py --gpus N --steps K --warmup W [--workload cfg2] [--impl ours|reference] [--scaling auto|strong|weak]

A *step* is one pass of the hot path over one batch of synthetic 16 kHz audio.
The workload is BASELINE.json configs[1] (cfg2: Conv-TasNet N=512,H=512,P=3,X=8,R=3,
ONE batch of 64 x 4 s).  Utterances are independent, so ranks shard them with no
data-path collective.  With N > 1 ranks the headline `value` is STRONG scaling - the
same 64-utterance batch cut into contiguous slices of 64/N per rank (BASELINE configs[1],
SURVEY.md 8e) - and the weak-scaling figure (64 utterances on every rank) is measured in
the same run and reported under the `weak` key.  cfg5 (streams per GPU) is weak by
definition.  One JSON line is printed by rank 0 (see README / DESIGN.md for the keys).

`--impl reference` times the UNMODIFIED reference (baseline/_ref, installed by
__graft_entry__.build()) - its SoTaskWrapModule.inference on the host cores - on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
SCALING_OF = {"strong": "strong", "weak": "weak"}
WORKLOADS = {
    # name: (config, batch per GPU, samples, enroll samples, description)
    "cfg1": ("cfg1", 1, 64000, None, "Conv-TasNet NS (N=512,H=512,P=3,X=8,R=3) 1 x 4 s @16 kHz"),
    # the reading of (N=512, B=128, H=512) that honours B (128-wide residual stream, 512-wide blocks)
    "cfg1b": ("cfg1b", 64, 64000, None, "Conv-TasNet NS cfg-1b (encoder/residual width B=128, H=512, P=3, X=8, R=3) batch 64 x 4 s @16 kHz"),
    "cfg2": ("cfg2", 64, 64000, None, "Conv-TasNet NS (N=512,H=512,P=3,X=8,R=3) batch 64 x 4 s @16 kHz"),
    "cfg3": ("cfg3", 32, 160000, None, "DPRNN-LSTM (chunk 100, 6 blocks, H=128 bi) batch 32 x 10 s @16 kHz per GPU"),
    "cfg4": ("cfg4", 64, 64000, 96000, "TSE: STFT 512/128 + TCN (H=256, dvec 192) + speaker net, 64 x (4 s mix + 6 s enroll) per GPU"),
    # widening row: cfg4 with GatedTCN blocks as in the reference's tse_unet_tcn recipes
    "cfg4_gated": ("cfg4_gated", 64, 64000, 96000, "TSE: STFT 512/128 + GatedTCN stack (H=256, dvec 192) + GatedTCN speaker net, 64 x (4 s mix + 6 s enroll) per GPU"),
    # widening row: the reference's own STFT-domain TSE recipe (egs/tse/model.py:184-243)
    "tse_unet_tcn_v0": ("tse_unet_tcn_v0", 64, 64000, 96000, "TSE recipe tse_unet_tcn_v0: STFT 512/128 + 6-layer U-Net shell + 15 GatedTCN blocks + GatedTCN speaker net, 64 x (4 s mix + 6 s enroll) per GPU"),
    # widening row: the egs/ns noise-suppression recipe (egs/ns/model.py:84-126)
    "ns_dpcrn_v0": ("ns_dpcrn_v0", 64, 64000, None, "NS recipe ns_dpcrn_v0: STFT 512/128 + 5-layer U-Net shell + 2 DPRNN-2D blocks (H=128), complex mask, 64 x 4 s per GPU"),
    # widening rows: the reference's demo recipe and its mel-front-end variant (egs/tse/model.py:418-463,509-558)
    "tse_skim_v0_causal": ("tse_skim_v0_causal", 32, 160000, 96000, "TSE recipe tse_skim_v0_causal: learned encoder 32/16 + causal SkiM (H=256, 4 blocks, seg 150, FiLM) + TCN speaker net, 32 x (10 s mix + 6 s enroll) per GPU"),
    "tse_skim_v2_causal": ("tse_skim_v2_causal", 32, 160000, 96000, "TSE recipe tse_skim_v2_causal: learned encoder 32/16 + causal SkiM (H=256) + mel front-end (FbankEnc 80 bands, SpecAugment) + TCN speaker net, 32 x (10 s mix + 6 s enroll) per GPU"),
    "ns_dparn_v0": ("ns_dparn_v0", 64, 64000, None, "NS recipe ns_dparn_v0: STFT 512/128 + 5-layer U-Net shell + 2 DPARN-2D blocks (8-head attention over frequency, LSTM over time), complex mask, 64 x 4 s per GPU"),
    # streaming: a step is one 10 ms hop of every stream (batch = concurrent streams, samples = hop)
    "cfg5": ("cfg5", 256, 160, None, "causal cLN Conv-TasNet (N=512,H=512,X=8,R=3), 10 ms hop, 256 concurrent streams per GPU, frame-by-frame"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs ~0.1-0.3 s to deliver its first sample, so the sampler is started BEFORE the warm-up (start()) and only
    the samples received between window_open() and window_close() are summarised.  A timed region shorter than a sampling
    period (8 ranks x 20 steps of 5.5 ms) can still end with no sample: the caller then keeps the same load running under
    the open window until two samples have arrived (timed_resident), and the summary says so (`extended`)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t_open, self.t_close, self.extended = None, None, False

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window_open(self):
        self.t_open = time.perf_counter()

    def window_close(self):
        self.t_close = time.perf_counter()

    def in_window(self):
        t1 = self.t_close if self.t_close is not None else float("inf")
        return [r for t, r in list(self.rows) if self.t_open is not None and self.t_open <= t <= t1 and len(r) >= 8]

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    # context-manager form: the window is the body
    def __enter__(self):
        self.start()
        self.window_open()
        return self

    def __exit__(self, *a):
        self.window_close()
        self.stop()

    def summary(self):
        rows = self.in_window()
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if self.extended:
            out["extended"] = "the timed region was shorter than the sampling period: the same steps kept running (untimed) until two samples had arrived"
        return out


def build_inputs(workload: str, rank: int, batch=None, span=None):
    """Seeded synthetic noisy speech.  span=(a, b): rows [a, b) of the ONE global batch (strong scaling: every rank derives
    its slice from the same seeded batch); otherwise a per-rank batch (weak scaling)."""
    from puresound_b200 import testing

    cfg, n, L, Le, _ = WORKLOADS[workload]
    n = batch or n
    if span is not None:
        mix, _ = testing.noisy_speech(n, L, seed=1234)
        enr = testing.noisy_speech(n, Le, seed=4321)[0] if Le else None
        a, b = span
        return mix[a:b].contiguous(), (enr[a:b].contiguous() if enr is not None else None)
    mix, _ = testing.noisy_speech(n, L, seed=1234 + rank)
    enr = testing.noisy_speech(n, Le, seed=4321 + rank)[0] if Le else None
    return mix, enr


def build_model(workload: str):
    from puresound_b200 import recipes, testing

    torch.manual_seed(0)
    m = recipes.baseline_config(WORKLOADS[workload][0]).eval()
    testing.perturb_(m, seed=1)
    return m


def ncu_traffic(kernel: str, shape, residual_frac: float = 0.0):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/traffic.json names the report it was read from); None when no capture of this shape is committed.
    residual_frac: share of the group's launches that read a residual - the figure is then the mean of the captured launch
    without and with one (like the group's algorithmic bytes)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as fh:
        table = json.load(fh)
    for key, t in table.items():
        if key.split("@")[0] == kernel and list(t.get("shape", [t.get("rows"), t.get("M"), t.get("K")])) == list(shape):
            base, with_res = t["dram_bytes_per_launch"], t.get("dram_bytes_per_launch_with_residual")
            if residual_frac > 0.0 and with_res:
                return (1.0 - residual_frac) * base + residual_frac * with_res
            return base
    return None


def kernel_rooflines(events, eager_ms, hbm_peak, tf_peak, peak_kind):
    """Group the live CUDA-event pairs of the eager pass by (kernel kind, shape) and put each group against the roofline
    that bounds it.  Algorithmic work per launch (DESIGN.md section 4 / SURVEY.md 8d):
      gemm   [rows, K] x [M, K]^T : 2*rows*M*K FLOP (3 tensor passes are ISSUED for the 3xBF16 split; frac counts useful
                                    FLOPs, so 1/3 is its ceiling), bytes 4*(rows*K + rows*M + M*K) + the residual / mask
                                    operand reads of the launches that have one (mean over the group)
      lstm   n_seq x L x D, H (+ fused input projection K_in): 2*4H*(H + K_in) FLOP per step, sequence and direction (3 passes)
      dwconv [B, T, C]            : 2*4*B*T*C bytes (one read, one write)
    Returns the groups sorted by their share of the step; the first is the dominant kernel."""
    groups, extras = {}, {}
    for kind, a, b, shape, extra in events:
        groups.setdefault((kind, tuple(shape)), []).append(a.elapsed_time(b))
        extras[(kind, tuple(shape))] = extras.get((kind, tuple(shape)), 0.0) + float(extra)
    out = []
    for (kind, shape), durs in groups.items():
        avg_ms = sum(durs) / len(durs)
        extra_bytes = extras[(kind, shape)] / len(durs)  # mean per launch: residual / mask operand reads of the group's GEMMs
        r = {"kernel": None, "shape": list(shape), "avg_launch_ms": avg_ms, "launches_timed": len(durs), "share_of_step": sum(durs) / eager_ms}
        if kind == "gemm":
            rows, M, K = shape
            flops, nbytes = 2.0 * rows * M * K, 4.0 * (rows * K + rows * M + M * K) + extra_bytes
            t_tensor, t_hbm = 3 * flops / (tf_peak * 1e12), nbytes / (hbm_peak * 1e9)
            r["kernel"] = "ps_gemm (1x1 conv / linear, %d x %d x %d)" % (rows, M, K)
            if t_tensor >= t_hbm:  # tensor-bound shape
                ach = flops / (avg_ms / 1e3) / 1e12
                r.update({"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "tensor_passes": 3,
                          "issued_frac": 3 * ach / tf_peak, "algorithmic_bytes": nbytes, "hbm_frac_if_memory_bound": nbytes / (avg_ms / 1e3) / 1e9 / hbm_peak})
            else:
                ach = nbytes / (avg_ms / 1e3) / 1e9
                r.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "algorithmic_bytes": nbytes})
            r["traffic"] = ncu_traffic("gemm", shape, min(1.0, extra_bytes / (4.0 * rows * M)))
        elif kind == "lstm":
            n_seq, L, H, D, K_in = shape
            flops = 2.0 * 4 * H * (H + K_in) * n_seq * L * D
            ach = flops / (avg_ms / 1e3) / 1e12
            r.update({"kernel": "ps_lstm (recurrence%s, %d seq x %d steps x %d dir, H=%d)" % (" + input projection" if K_in else "", n_seq, L, D, H),
                      "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "tensor_passes": 3,
                      "issued_frac": 3 * ach / tf_peak, "step_latency_us": 1e3 * avg_ms / L,
                      "note": "a recurrence: L dependent steps per launch, so latency per step (step_latency_us) bounds it before the tensor peak does",
                      "traffic": ncu_traffic("lstm", shape)})
        elif kind == "dwconv":
            B, T, Cn = shape
            nbytes = 2.0 * 4 * B * T * Cn
            ach = nbytes / (avg_ms / 1e3) / 1e9
            r.update({"kernel": "ps_dwconv (dilated depthwise conv + norm/PReLU prologue + Welford partials, %d x %d x %d)" % (B, T, Cn),
                      "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "algorithmic_bytes": nbytes,
                      "traffic": ncu_traffic("dwconv", shape)})
        r["peak_source"] = f"{peak_kind} ({'bf16 sustained' if r.get('bound') == 'tensor' else 'copy bandwidth'})"
        out.append(r)
    out.sort(key=lambda r: -r["share_of_step"])
    return out


def streaming_bench(args, rank, world, local_rank):
    """cfg5: frame-by-frame streaming.  value = audio-s/s over all streams; also per-hop latency p50/p99 at S=256 and S=1."""
    from puresound_b200 import ops, recipes, sharding, testing
    from puresound_b200.streaming.conv_tasnet_inference import StreamingConvTasNet, StreamingSeparator
    from puresound_b200.nnet.base_nn import SoTaskWrapModule
    from puresound_b200.nnet.lobe.encoder import FreeEncDec

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    ops.require_device()
    torch.manual_seed(0)
    m = SoTaskWrapModule(FreeEncDec(320, 512, 160),
                         StreamingConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8,
                                             tcn_norm="cLN", dconv_norm="cLN", causal=True), mask_constraint="ReLU", verbose=False).eval()
    testing.perturb_(m, seed=1)
    m = m.to(dev)
    _, S, hop, _, desc = WORKLOADS["cfg5"]

    def run(n_streams, steps, warm):
        sep = StreamingSeparator(m, use_graph=True)
        sep.init_status(n_streams)
        chunk_h = testing.white(n_streams, hop, amp=0.1, seed=99 + rank).pin_memory()
        chunk_d = chunk_h.to(dev)
        for _ in range(warm + 2):
            sep.step_wave(chunk_d)
        torch.cuda.synchronize()
        lat = []
        l0 = ops.launch_count
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sep.step_wave(chunk_d)
            e1.record()
            e1.synchronize()
            lat.append(e0.elapsed_time(e1))
        launches = ops.launch_count - l0
        t0 = time.perf_counter()
        for _ in range(steps):
            out = sep.step_wave(chunk_h)  # host chunk in, host chunk out
        e2e = (time.perf_counter() - t0) * 1e3 / steps
        lat.sort()
        return lat, e2e, launches, out

    # kernels of one hop (a graph replay launches them without passing through ops.launch_count): counted on an eager hop
    sep_e = StreamingSeparator(m, use_graph=False)
    sep_e.init_status(S)
    probe = testing.white(S, hop, amp=0.1, seed=5).to(dev)
    for _ in range(3):
        sep_e.step_wave(probe)
    l0 = ops.launch_count
    sep_e.step_wave(probe)
    kernels_per_hop = ops.launch_count - l0
    hop_kernel = sep_e._hop is not None
    del sep_e
    steps = max(args.steps, 50)
    sharding.barrier(dev)
    with ClockSampler(local_rank) as clk:
        lat256, e2e256, launches, out = run(S, steps, max(args.warmup, 3))
    lat1, e2e1, _, _ = run(1, steps, max(args.warmup, 3))
    lat16, _, _, _ = run(16, steps, max(args.warmup, 3))
    ms = sharding.max_over_ranks(sum(lat256) / len(lat256), dev)
    e2e_ms = sharding.max_over_ranks(e2e256, dev)
    audio_s = S * hop / SR
    # roofline of a hop: every weight is read once per hop (all S streams share it) and every stream
    # touches (P-1) taps + one new slot of each block's dilation-history ring - memory traffic that has to move whatever
    # the kernel structure is; the GEMM FLOPs (2 * 19.24 MMAC per stream-hop) are 3 passes of 9.85 GFLOP = 21 us of tensor time
    hbm_peak, tf_peak, peak_kind = peaks()
    w_bytes = 4 * sum(p.numel() for p in m.parameters())
    blocks = [b for st in m.masker.tcn_list for b in st]
    state_bytes = 4 * S * sum(b.kernel * b.hid_channels for b in blocks)
    hop_bytes = float(w_bytes + state_bytes)
    what = ("ONE persistent cooperative kernel (ps_stream_hop: ~100 phases behind grid barriers)" if hop_kernel
            else f"a CUDA-graph replay of {kernels_per_hop} kernels")
    launches = kernels_per_hop * steps
    roof = {"bound": "hbm", "kernel": f"one hop of {S} streams (encoder, 24 TCN blocks, decoder) = {what}",
            "achieved": hop_bytes / (ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": hop_bytes / (ms / 1e3) / 1e9 / hbm_peak,
            "traffic": None, "algorithmic_bytes": hop_bytes, "weights_bytes": w_bytes, "state_bytes": state_bytes,
            "peak_source": f"{peak_kind} (copy bandwidth)", "avg_launch_ms": ms, "launches_timed": steps,
            "note": "latency-bound: a hop is ~100 dependent phases (each ~4 us of loads -> compute -> stores, then a ~2 us grid barrier); frac is the hop's unavoidable bytes against the HBM peak"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        a_s, times, kind = cpu_reference_run("cfg5", 4, 2, 1, os.cpu_count() or 1)
        cpu = {"value": a_s / min(times), "unit": "audio-s/s", "cores": os.cpu_count() or 1, "kind": kind,
               "sample": "offline causal forward of 4 x 4 s utterances (the reference has no frame-by-frame Conv-TasNet), best of 2 after 1 warm-up"}
    if rank == 0:
        pct = lambda v, q: v[min(len(v) - 1, int(q * len(v)))]
        print(json.dumps({
            "metric": "audio-sec/sec", "value": world * audio_s / (ms / 1e3), "unit": "audio-s/s", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"cfg5: {desc}", "graph": "one CUDA graph replay per hop", "kernels_per_hop": kernels_per_hop, "hop_budget_ms": 10.0,
                                            "l2": "per-hop state touch (815 MB of ring buffers at S=256) exceeds L2"},
            "latency_ms": {"S=256": {"p50": pct(lat256, 0.5), "p99": pct(lat256, 0.99)}, "S=1": {"p50": pct(lat1, 0.5), "p99": pct(lat1, 0.99)},
                           "S=16": {"p50": pct(lat16, 0.5), "p99": pct(lat16, 0.99)}, "e2e_host_S=1": e2e1},
            "clocks": clk.summary(),
            "e2e": {"value": world * audio_s / (e2e_ms / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": S * hop * 4, "d2h_bytes_per_step": S * hop * 4,
                    "ms_per_step": e2e_ms},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        }))
    return 0


def cpu_reference_run(workload: str, sample_batch: int, steps: int, warmup: int, threads: int):
    """The reference's own CPU forward of the path: the UNMODIFIED reference installed in baseline/_ref
    (`SoTaskWrapModule.inference`, puresound/nnet/base_nn.py:690-722, built from the reference's classes by
    baseline/ref_models.py with the same seeded weights as the B200 arm) on all host threads.  Where baseline/_ref is
    absent (never on a box the snapshot was pushed from the authoring container) the oracle port is timed instead.
    Returns (audio seconds per step, step times, kind)."""
    torch.set_num_threads(threads)
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_models

    cfg, _, L, Le, _ = WORKLOADS[workload]
    if workload == "cfg5":
        L = 4 * SR  # the reference has no frame-by-frame Conv-TasNet: its arithmetic for cfg5 is the offline causal forward
    from puresound_b200 import testing

    mix, _ = testing.noisy_speech(sample_batch, L, seed=1234)
    enr = testing.noisy_speech(sample_batch, Le, seed=4321)[0] if Le else None
    if ref_models.available():
        m = ref_models.build(cfg)
        kind = "reference"

        def run():
            with torch.no_grad():
                return m.inference(mix, enr) if enr is not None else m.inference(mix)
    else:
        from oracle import describe as D
        from oracle import separator_ref as R

        om = build_model(workload)
        sd, dcfg = om.state_dict(), D.describe(om)
        kind = "port"

        def run():
            return R.inference(sd, dcfg, mix, enr)
    for _ in range(warmup):
        run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        y = run()
        times.append(time.perf_counter() - t0)
    assert torch.isfinite(y).all()
    return sample_batch * L / SR, times, kind


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="utterances in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="N > 1: strong = ONE batch cut into slices (BASELINE configs[1]); weak = a full batch on every rank; auto = strong, with the weak figure under the `weak` key")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling measurement of a strong-scaling run")
    ap.add_argument("--gemm-backend", default="auto", choices=["auto", "simt"],
                    help="auto: tcgen05 3xBF16 GEMM where eligible, exact-fp32 CUDA-core GEMM elsewhere; simt: CUDA-core everywhere")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name, batch, L, Le, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample_batch = args.cpu_sample_batch or {"cfg1": 1, "cfg1b": 8, "cfg2": 4, "cfg3": 2, "cfg4": 8, "cfg4_gated": 4, "tse_unet_tcn_v0": 2, "ns_dpcrn_v0": 4, "ns_dparn_v0": 4, "cfg5": 4,
                                                "tse_skim_v0_causal": 2, "tse_skim_v2_causal": 2}[args.workload]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = max(1, args.steps), max(1, args.warmup)
        audio_s, times, kind = cpu_reference_run(args.workload, sample_batch, steps, warm, cores)
        t = sum(times) / len(times)
        v = audio_s / t
        if args.workload == "cfg5":
            sample = f"offline causal forward of {sample_batch} x 4 s utterances per step (the reference has no frame-by-frame Conv-TasNet), {cores} torch threads"
        else:
            sample = f"{sample_batch} x {L / SR:.0f} s utterances of the {batch}-utterance batch per step, {cores} torch threads"
        what = "unmodified reference (baseline/_ref), SoTaskWrapModule.inference on CPU" if kind == "reference" else "oracle port (baseline/_ref absent)"
        print(json.dumps({
            "impl": "reference", "metric": "audio-sec/sec", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": SCALING_OF.get(args.scaling, "weak"), "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "sample": sample, "what": what},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    if args.workload == "cfg5":
        return streaming_bench(args, rank, world, local_rank)

    # ------------------------------------------------------------------ our arm (B200)
    from puresound_b200 import ops, sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL carries no data of the path: it times it (a barrier around the timed region, one scalar MAX of the device time)
        dist.init_process_group("nccl", device_id=dev)
    ops.require_device()
    if args.gemm_backend != "auto":
        ops.force_gemm_backend = ops.GEMM_SIMT

    scaling = args.scaling if args.scaling != "auto" else ("strong" if world > 1 else "weak")
    if scaling == "strong" and batch < world:
        scaling = "weak"  # nothing to split (cfg1: one utterance)
    model = build_model(args.workload).to(dev)

    def barrier():
        sharding.barrier(dev)

    def max_over_ranks(ms: float) -> float:
        return sharding.max_over_ranks(ms, dev)

    def timed_resident(mix_d, enr_d, steps, warm):
        """K device-timed replays of the public API on resident inputs; returns (ms for the K steps, clock summary, last result)."""
        clk = ClockSampler(local_rank).start()  # (nvidia-smi delivers its first sample after ~0.2 s: started before the warm-up)
        with torch.no_grad():
            for _ in range(warm):
                y = model.inference(mix_d, enr_d)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            clk.window_open()
            e0.record()
            for _ in range(steps):
                y = model.inference(mix_d, enr_d)  # public API, device tensors in and out (CUDA-graph replay)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            # a region shorter than the sampling period: keep the same load running (untimed) until two samples arrived
            t_give_up = time.perf_counter() + 1.5
            while len(clk.in_window()) < 2 and clk.proc is not None and time.perf_counter() < t_give_up:
                clk.extended = True
                for _ in range(max(1, steps // 4)):
                    y = model.inference(mix_d, enr_d)
                torch.cuda.synchronize(dev)
            clk.window_close()
            clk.stop()
            barrier()
        return max_over_ranks(ms), clk.summary(), y

    # the rank's share of the job
    if scaling == "strong":
        a, b = sharding.shard_slice(batch, rank, world)
        mix_h, enr_h = build_inputs(args.workload, rank, span=(a, b))
        job_audio_s = batch * L / SR           # ONE batch for the whole job
    else:
        mix_h, enr_h = build_inputs(args.workload, rank)
        job_audio_s = world * batch * L / SR   # a full batch per rank
    mix_h = mix_h.pin_memory()
    enr_h = enr_h.pin_memory() if enr_h is not None else None
    mix_d = mix_h.to(dev)
    enr_d = enr_h.to(dev) if enr_h is not None else None
    warm = max(args.warmup, 3)

    ms, clocks, y = timed_resident(mix_d, enr_d, args.steps, warm)
    with torch.no_grad():
        # ---- roofline pass: the same K steps launched eagerly (no graph) with a CUDA-event pair around every GEMM / LSTM /
        #      depthwise-conv launch on the launching stream, which also counts the kernels of a step ----
        graphed, model.use_cuda_graph = model.use_cuda_graph, False
        ops.kernel_events = []
        launches0 = ops.launch_count
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            y = model.inference(mix_d, enr_d)
        r1.record()
        barrier()
        launches = ops.launch_count - launches0
        eager_ms = r0.elapsed_time(r1)
        events, ops.kernel_events = ops.kernel_events, None
        model.use_cuda_graph = graphed
        # ---- e2e: the public API with HOST buffers; H2D of the inputs and D2H of the result inside the timed region ----
        for _ in range(2):
            yh = model.inference(mix_h, enr_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            yh = model.inference(mix_h, enr_h)
        torch.cuda.synchronize()
        e2e_seq_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        # the same K host batches through the serving loop of the public API (inference_stream): every step still copies its
        # inputs from pinned host memory and its result back to the host, but batch i+1's H2D and result i-1's D2H run on
        # their own streams under batch i's forward
        # (six warm-up batches: the loop keeps up to four pinned result buffers in flight and torch's caching host allocator
        # only reaches that steady state after a few batches - a fresh cudaHostAlloc is a device-wide synchronisation)
        for yh in model.inference_stream([(mix_h, enr_h)] * 6, reuse_host_buffers=True):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for yh in model.inference_stream(((mix_h, enr_h) for _ in range(args.steps)), reuse_host_buffers=True):
            n_out += 1  # (results land in a ring of pinned buffers: a fresh pinned allocation per batch is a device-wide stall)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        assert n_out == args.steps
    assert torch.isfinite(yh).all()

    value = job_audio_s * args.steps / (ms / 1e3)
    e2e_value = job_audio_s * args.steps / (e2e_ms / 1e3)
    h2d = mix_h.numel() * 4 + (enr_h.numel() * 4 if enr_h is not None else 0)
    d2h = yh.numel() * 4

    # ---- weak-scaling figure beside the strong one (N > 1): a full batch on every rank, same timing rules ----
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        wm, we = build_inputs(args.workload, rank)
        wm_d, we_d = wm.to(dev), (we.to(dev) if we is not None else None)
        wms, _, _ = timed_resident(wm_d, we_d, args.steps, warm)
        weak = {"value": world * batch * L / SR * args.steps / (wms / 1e3), "unit": "audio-s/s", "ms_per_step": wms / args.steps,
                "batch_per_rank": batch, "scaling": "weak"}
        del wm_d, we_d

    hbm_peak, tf_peak, peak_kind = peaks()
    roofs = kernel_rooflines(events, eager_ms, hbm_peak, tf_peak, peak_kind) if events else []
    roof = None
    if roofs:
        roof = dict(roofs[0])
        roof["eager_ms_per_step"] = eager_ms / args.steps
        if roof.get("bound") == "tensor":
            roof["note"] = ("3xBF16 split: useful FLOPs / measured bf16 peak (three tensor passes are issued, so 1/3 is the ceiling of frac); "
                            "the step runs at the 1000 W power cap (see clocks.reasons)" + ("; " + roof["note"] if roof.get("note") else ""))
        roof["other_kernels"] = [{k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "launches_timed", "share_of_step", "algorithmic_bytes", "traffic", "step_latency_us") if k in r}
                                 for r in roofs[1:4]]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        audio_s, times, kind = cpu_reference_run(args.workload, sample_batch, 2, 1, cores)
        cpu = {"value": audio_s / min(times), "unit": "audio-s/s", "cores": cores, "kind": kind,
               "sample": f"{sample_batch} x {L / SR:.0f} s utterances of the {batch}-utterance batch, best of 2 after 1 warm-up"}

    if rank == 0:
        per_rank = mix_h.shape[0]
        print(json.dumps({
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}", "l2": "activations (524 MB per tensor at cfg2) exceed the 126 MB L2",
                                            "gemm_backend": args.gemm_backend, "accumulate": "fp32",
                                            "sharding": (f"one batch of {batch} utterances cut into contiguous slices, {per_rank} on rank 0; no data-path collective "
                                                         "(NCCL only for the timing barrier and the MAX of the device time)") if scaling == "strong"
                                            else f"{batch} utterances on every rank; no data-path collective (NCCL only for the timing barrier and the MAX of the device time)",
                                            "timed_region": "model.inference(device tensors): CUDA-graph replay of the whole forward",
                                            "roofline_pass": "the same K steps re-run eagerly with a CUDA-event pair around every GEMM / LSTM / depthwise-conv launch"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "ms_per_step": e2e_ms / args.steps, "api": "model.inference_stream(host batches, reuse_host_buffers=True): H2D / forward / D2H on three streams, results in a ring of pinned buffers",
                                      "sequential_api_ms_per_step": e2e_seq_ms / args.steps,
                                      "sequential_api_value": job_audio_s * args.steps / (e2e_seq_ms / 1e3)},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "weak": weak,
        }))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
