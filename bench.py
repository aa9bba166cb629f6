#!/usr/bin/env python
"""Benchmark of the BayesVLM post-hoc Laplace hot path on B200 (contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline metric (BASELINE.json): predictive pairs/s on the CLIP ViT-L-14 ImageNet-shaped workload -- 50 000 images x
1000 classes, D=768, d_img=1024, d_txt=768, logit mean + variance -- per GPU (weak scaling: every rank owns its own
50k-image shard, class side replicated, no data-path collective).  A "step" is one pass of the predictive over the
rank's shard.  The same line carries:
  roofline       live CUDA-event time of the dominant tensor-core kernel against MEASURED_PEAKS.json
  parity_check   384 random rows of the timed output against the fp64 oracle (the line FAILS on a breach)
  e2e            the same metric through CLIP.predict_host with HOST buffers (H2D + kernels + D2H inside the timed region)
  with_probs     the step with the probit softmax as third output (north-star config 3)
  kfac           config 2 (ViT-B-32, class batches of 32768): image-modality factors, weak scaling, + SYRK roofline entry
  kfac_cfg2      config 2 in full: 32 class batches, BOTH modalities, sharded over the ranks + one all-reduce each (strong)
  kfac_h14 / kfac_siglip   configs 4 / 5 estimation shapes
  epig           config 5: pool 100 000 / N rows per rank x 10 000 targets (Cl = 10, and Cl = 65 on a smaller pool), the
                 reductions' achieved GB/s and the exact-match rate against the reference sequence on the same GPU
  cpu_baseline   the reference algorithm (oracle/torch_port.py) on the box's host cores, same workload, bounded sample

`--impl reference` times the reference's own algorithm (torch CPU restatement in oracle/torch_port.py -- the reference
is a PyTorch program that cannot travel to the GPU box) on the FULL 50 000 x 1000 call.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LS = math.log(100.0)
PRED = dict(name="clip-vit-l-14 imagenet-1k-shaped predictive", N=50_000, C=1000, D=768, d_img=1024, d_txt=768,
            lam_img=605.255, lam_txt=220.124, seed=3001)
EPIG = dict(name="siglip-shaped EPIG scoring (config 5): pool x target, K=100 MC samples, chunk 4096", pool=100_000,
            pool_cl65=16_384, target=10_000, K=100, chunk=4096, seed=5001)
KFAC = dict(name="clip-vit-b-32 kfac (InfoNCE), class batches of 32768", num_classes=32768, batch_size=5, D=512, d_img=768,
            d_txt=512, seed=2001, total_class_batches=32)
KFAC_H14 = dict(name="clip-vit-h-14 kfac (InfoNCE), class batches of 32768", num_classes=32768, batch_size=5, D=1024,
                d_img=1280, d_txt=1024, seed=4001)
KFAC_SIGLIP = dict(name="siglip-base kfac (sigmoid loss), class batches of 32768", num_classes=32768, batch_size=5, D=768,
                   d_img=3072, d_txt=768, seed=5002, logit_scale=4.765, logit_bias=-12.93)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# synthetic inputs: seeded CPU generators, fp32
# ----------------------------------------------------------------------------------------------------------------------
def surrogate_spd(gen, d, scale):
    w = torch.randn(4 * d, d, generator=gen, dtype=torch.float32).double()
    return ((w.T @ w) / math.sqrt(4 * d) * scale).float()


def predictive_inputs(cfg, rank, n_rows=None):
    gen = torch.Generator().manual_seed(cfg["seed"])
    A_img, A_txt = surrogate_spd(gen, cfg["d_img"], 3e3), surrogate_spd(gen, cfg["d_txt"], 3e3)
    B_img, B_txt = surrogate_spd(gen, cfg["D"], 20.0), surrogate_spd(gen, cfg["D"], 20.0)
    txt_e = torch.randn(cfg["C"], cfg["D"], generator=gen)
    txt_a = torch.randn(cfg["C"], cfg["d_txt"], generator=gen)
    gen_r = torch.Generator().manual_seed(cfg["seed"] + 17 * (rank + 1))
    n = cfg["N"] if n_rows is None else n_rows
    img_e = torch.randn(n, cfg["D"], generator=gen_r)
    img_a = torch.randn(n, cfg["d_img"], generator=gen_r)
    return dict(A_img=A_img, A_txt=A_txt, B_img=B_img, B_txt=B_txt, txt_e=txt_e, txt_a=txt_a, img_e=img_e, img_a=img_a)


def covariances(t, cfg, device):
    def inv(F, lam):
        F = F.to(device).double()
        return torch.linalg.inv(F + math.sqrt(lam) * torch.eye(F.shape[0], dtype=torch.float64, device=device)).float()

    return (inv(t["A_img"], cfg["lam_img"]), inv(t["B_img"], cfg["lam_img"]), inv(t["A_txt"], cfg["lam_txt"]),
            inv(t["B_txt"], cfg["lam_txt"]))


def kfac_inputs(cfg, n, seed, device="cpu", with_txt_act=False):
    """LAION-shaped pairs: emb_img = z + 1.5 eps, emb_txt = z + 1.5 eps' (paired cosine ~0.31); activations N(0,1)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=gen, device=device)
    z = rn(n, cfg["D"])
    out = (z + 1.5 * rn(n, cfg["D"]), z + 1.5 * rn(n, cfg["D"]), rn(n, cfg["d_img"]))
    if with_txt_act:
        out = out + (rn(n, cfg["d_txt"]),)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed regions, in-process through NVML (no nvidia-smi fork per sample)
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, dev: torch.device, period=0.05):
        self.samples, self._stop, self._thr, self.period, self.handle, self.nv = [], threading.Event(), None, period, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(dev)
            bus_id = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover - NVML missing
            log("clock sampling unavailable:", repr(exc))

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for n, bit in self.REASONS.items() if any(s[1] & bit for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "how": "NVML in-process, every 50 ms during the timed regions"}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------------
# reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_predictive_rate(cfg, rows, steps, warmup):
    from oracle import torch_port as T

    torch.set_num_threads(os.cpu_count() or 1)
    t = predictive_inputs(cfg, 0, n_rows=rows)
    Ai, Bi, At, Bt = covariances(t, cfg, "cpu")
    args = (t["img_e"], t["img_a"], t["txt_e"], t["txt_a"], Ai, Bi, At, Bt, LS)
    for _ in range(warmup):
        T.predictive(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        T.predictive(*args)
    dt = (time.perf_counter() - t0) / steps
    return rows * cfg["C"] / dt, dt


def cpu_kfac_rate(cfg, data_batches):
    from oracle import torch_port as T

    torch.set_num_threads(os.cpu_count() or 1)
    n = cfg["num_classes"]
    e_img, e_txt, a_img = kfac_inputs(cfg, n, cfg["seed"])
    T.kfac_ggn(e_img, a_img, e_txt, n, cfg["batch_size"], LS, max_data_batches=1)  # warm-up
    t0 = time.perf_counter()
    T.kfac_ggn(e_img, a_img, e_txt, n, cfg["batch_size"], LS, max_data_batches=data_batches)
    dt = time.perf_counter() - t0
    return data_batches * cfg["batch_size"] / dt, dt


def gpu_eager_rates(cfg, kc, dev):
    """The reference algorithm in stock torch eager ON THE SAME B200 (true fp32, no TF32): the tougher, informational baseline
    SURVEY.md section 8(d) asks for next to the host-core one.  Predictive: the full 50k x 1000 step; KFAC: the reference's
    double loop over 8 data batches of 5 sources against one class batch of 32768 targets."""
    from oracle import torch_port as T

    t = predictive_inputs(cfg, 0)
    Ai, Bi, At, Bt = covariances(t, cfg, dev)
    a = tuple(t[k].to(dev) for k in ("img_e", "img_a", "txt_e", "txt_a")) + (Ai, Bi, At, Bt, LS)
    for _ in range(2):
        T.predictive(*a)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        T.predictive(*a)
    torch.cuda.synchronize(dev)
    pdt = (time.perf_counter() - t0) / 5
    n = kc["num_classes"]
    e_img, e_txt, a_img = kfac_inputs(kc, n, kc["seed"], device=dev)
    T.kfac_ggn(e_img, a_img, e_txt, n, kc["batch_size"], LS, max_data_batches=1)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    T.kfac_ggn(e_img, a_img, e_txt, n, kc["batch_size"], LS, max_data_batches=8)
    torch.cuda.synchronize(dev)
    kdt = time.perf_counter() - t0
    return {"predictive_pairs_per_s": cfg["N"] * cfg["C"] / pdt, "predictive_ms_per_step": pdt * 1e3,
            "kfac_samples_per_s": 8 * kc["batch_size"] / kdt,
            "what": "oracle/torch_port (the reference's ATen operation sequence) in torch eager fp32 on the same GPU, inputs resident"}


def pred_config(cfg, world, extra=None):
    c = {"workload": cfg["name"], "images_per_gpu": cfg["N"], "classes": cfg["C"], "D": cfg["D"], "d_img": cfg["d_img"],
         "d_txt": cfg["d_txt"]}
    if extra:
        c.update(extra)
    return c


def run_reference(args, rank):
    """The reference arm: the reference's own algorithm for the headline path on the host cores, on the SAME config as the
    B200 arm -- one step = one full 50 000 x 1000 `CLIP.forward`-equivalent call (oracle/torch_port.predictive, ~1 s).  The
    number of timed calls is capped so that the run ends within a few minutes; the line reports what was actually timed."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    rate, dt = cpu_predictive_rate(PRED, PRED["N"], steps, warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "predictive_pairs_per_s", "value": rate, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": pred_config(PRED, 1),
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full {PRED['N']} x {PRED['C']} calls of the reference algorithm (torch CPU fp32, "
                                   f"oracle/torch_port.predictive; {dt * 1e3:.0f} ms per call, {cores} threads); requested "
                                   f"steps={args.steps} capped at 20 to bound the run"},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch.distributed as dist

    from bayesvlm_b200 import _lib
    from bayesvlm_b200.hessians import KroneckerFactorizedCovariance as KFC
    from bayesvlm_b200.hessians import kfac_ggn
    from bayesvlm_b200.vlm import CLIP, SIGLIP, EncoderResult

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); bayesvlm_b200 has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.lib.bvlm_device_check(), "bvlm_device_check")
    peaks, peaks_src = measured_peaks()
    peak = peaks["bf16_tflops_sustained"]
    peak_burst = peaks.get("bf16_tflops", peak)
    use_dist = world > 1

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, settle=2):
        """`settle` untimed calls right after the barrier (part of the warm-up: the clocks are up again when the region
        starts), then exactly `steps` calls between two CUDA events; returns ms per call, max over ranks."""
        barrier_sync()
        for _ in range(settle):
            fn()
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        beg.record()
        for _ in range(steps):
            fn()
        end.record()
        barrier_sync()
        return max_over_ranks(beg.elapsed_time(end) / steps)

    clocks = ClockSampler(dev)

    # ------------------------------------------------------------------ predictive (headline)
    cfg = PRED
    t = predictive_inputs(cfg, rank)
    Ai, Bi, At, Bt = covariances(t, cfg, dev)
    model = CLIP(logit_scale=LS, device=dev) if args.precision is None else CLIP(logit_scale=LS, device=dev,
                                                                                 precision=args.precision)
    model.set_covariances(KFC(Ai, Bi), KFC(At, Bt))
    img = EncoderResult(t["img_e"].to(dev), t["img_a"].to(dev))
    txt = EncoderResult(t["txt_e"].to(dev), t["txt_a"].to(dev))
    W, K = max(3, args.warmup), max(1, args.steps)
    pairs = cfg["N"] * cfg["C"]
    state = {}

    def step():
        state["out"] = model(img, txt)

    with torch.no_grad():
        for _ in range(max(1, W - 2)):
            step()
        l0 = _lib.launch_count()
        clocks.__enter__()
        ms_step = timed(step, K, settle=2)
        launches = (_lib.launch_count() - l0) * K // (K + 2)
        out = state["out"]
        # ---- parity of the TIMED output: 384 random rows against the fp64 oracle (same tolerances as tests/)
        from oracle import laplace_oracle as O

        gsel = torch.Generator().manual_seed(1234 + rank)
        rows = torch.randperm(cfg["N"], generator=gsel)[:384]
        ref_m, ref_v = O.predictive(t["img_e"][rows].numpy(), t["img_a"][rows].numpy(), t["txt_e"].numpy(), t["txt_a"].numpy(),
                                    Ai.cpu().numpy(), Bi.cpu().numpy(), At.cpu().numpy(), Bt.cpu().numpy(), LS, dtype=np.float64)
        got_m, got_v = out.mean[rows.to(dev)].cpu().numpy(), out.var[rows.to(dev)].cpu().numpy()
        s_lin = math.exp(LS)
        ex_m = float((np.abs(got_m - ref_m) / (1e-3 * np.maximum(np.abs(ref_m), 0.01 * s_lin))).max())
        ex_v = float((np.abs(got_v - ref_v) / (1e-3 * np.abs(ref_v))).max())
        parity = {"rows": int(rows.numel()), "max_excess": max(ex_m, ex_v), "mean_excess": ex_m, "var_excess": ex_v,
                  "what": "max over 384 random rows x 1000 classes of |err| / tolerance (1e-3 max(|ref|, 0.01 s) for the mean, "
                          "1e-3 |ref| for the variance) against oracle.predictive in fp64; <= 1 passes"}
        ok = torch.tensor([1.0 if parity["max_excess"] <= 1.0 else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) < 0.5:
            raise RuntimeError(f"predictive parity check failed on the timed output: {parity}")
        # ---- second, shorter region with a CUDA event pair around every launch (the live roofline numerator); kept apart
        #      so that the event records do not sit inside the headline timing
        _lib.timing_enable(True)
        for _ in range(max(1, min(K, 20))):
            step()
        barrier_sync()
        _lib.timing_enable(False)
        kern = _lib.timing_collect()
        # ---- the same step with the probit softmax as third output (north-star config 3: "+ probit softmax")
        def step_probs():
            state["probs"] = model._compute_probabilistic_logits_smith(img, txt, return_probs=True)

        for _ in range(2):
            step_probs()
        ms_probs = timed(step_probs, max(1, min(K, 50)))
        pl, probs = state["probs"]
        ref_p = O.probit_softmax(ref_m, ref_v, dtype=np.float64)
        perr = float(np.abs(probs[rows.to(dev)].cpu().numpy() - ref_p).max())
        if perr > 1e-4:
            raise RuntimeError(f"probit softmax parity check failed: max abs {perr}")
        _lib.timing_enable(True)
        for _ in range(5):
            step_probs()
        barrier_sync()
        _lib.timing_enable(False)
        kern_p = _lib.timing_collect()
        del pl, probs
        state.pop("probs")
    value = world * pairs / (ms_step * 1e-3)
    with_probs = {"value": world * pairs / (ms_probs * 1e-3), "unit": "pairs/s", "ms_per_step": ms_probs,
                  "outputs": "mean + var + probit-softmax probabilities (fp32)", "probs_max_abs_err": perr,
                  "hbm_floor_ms": (12.0 * pairs + 4.0 * cfg["N"] * (cfg["D"] + cfg["d_img"])) / (peaks["hbm_gbs"] * 1e9) * 1e3,
                  "kernels": {k: {"launches": v[0], "avg_ms": v[1] / v[0]} for k, v in kern_p.items()}}

    # roofline of the dominant tensor-core kernel, timed live with CUDA events on its own stream
    algo_flops = {"predictive": 2.0 * cfg["N"] * cfg["C"] * cfg["D"],            # one [N,D]x[D,C] product (2D per pair)
                  "quadform": 2.0 * cfg["N"] * cfg["d_img"] ** 2}               # a^T A^-1 a as the reference counts it (2 d^2 per image)
    tensor_kernels = {k: v for k, v in kern.items() if k in algo_flops}
    dom = max(tensor_kernels, key=lambda k: tensor_kernels[k][1])
    avg_ms = kern[dom][1] / kern[dom][0]
    ach = algo_flops[dom] / (avg_ms * 1e-3) / 1e12
    try:
        traffic_rec = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        traffic, traffic_src = traffic_rec.get(dom), traffic_rec.get("_source")
    except Exception:
        traffic, traffic_src = None, None
    step_flops = algo_flops["predictive"] + algo_flops["quadform"]
    roofline = {"bound": "tensor", "kernel": f"gemm2_tn_kernel<{dom}> (CTA-pair tcgen05 engine)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "frac_of_burst_peak": ach / peak_burst, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks_src + ", sustained bf16", "avg_launch_ms": avg_ms, "launches": kern[dom][0],
                "kernels": {k: {"launches": v[0], "avg_ms": v[1] / v[0],
                                "algo_tflops": algo_flops.get(k, 0.0) / (v[1] / v[0] * 1e-3) / 1e12} for k, v in kern.items()},
                # the predictive sits on the ridge (report both): the same launch against the HBM roof, algorithmic
                # bytes = mean + var written (8 B per pair) + the operands read once (fp16 + fp8 pair or fp16 hi|lo: 4 B per element)
                "hbm": {"achieved": (8.0 * pairs + (2.0 if model.precision == "fp16" else 4.0) * (cfg["N"] + cfg["C"]) * cfg["D"]) / (kern["predictive"][1] / kern["predictive"][0] * 1e-3) / 1e9
                        if "predictive" in kern else None, "peak": peaks["hbm_gbs"], "unit": "GB/s"},
                # the same launch counted by the tensor work it EXECUTES (fp16-equivalent MMA passes over K: the error-compensated
                # modes run 2 (fp16 + two FP8 half-cost phases) or 3 (hi.hi + lo.hi + hi.lo) passes for one algorithmic product),
                # and by the bytes it moves through the L2-to-SM interface (operand tiles in + results out)
                "executed": {"mma_passes": {"fp16": 1, "fp16+fp8": 2, "fp16x3": 3}.get(model.precision, 1),
                             "tflops_fp16_equiv": ach * {"fp16": 1, "fp16+fp8": 2, "fp16x3": 3}.get(model.precision, 1) if dom == "predictive" else ach / 2,
                             "frac_of_sustained_peak": (ach * {"fp16": 1, "fp16+fp8": 2, "fp16x3": 3}.get(model.precision, 1) if dom == "predictive" else ach / 2) / peak,
                             "l2_to_sm_bytes": (math.ceil(cfg["N"] / 256) * math.ceil(cfg["C"] / 256) * 512 * cfg["D"] * {"fp16": 2, "fp16+fp8": 4, "fp16x3": 4}.get(model.precision, 2)
                                                + 8.0 * pairs) if dom == "predictive" else None,
                             "l2_to_sm_tbs": ((math.ceil(cfg["N"] / 256) * math.ceil(cfg["C"] / 256) * 512 * cfg["D"] * {"fp16": 2, "fp16+fp8": 4, "fp16x3": 4}.get(model.precision, 2)
                                               + 8.0 * pairs) / (avg_ms * 1e-3) / 1e12) if dom == "predictive" else None,
                             "l2_to_sm_ceiling_tbs": "~10-11 (6300 B/clk chip-wide at the 1.55-1.75 GHz these kernels run at)"},
                "step_algo_tflops": step_flops / (ms_step * 1e-3) / 1e12,
                "step_frac_of_sustained_peak": step_flops / (ms_step * 1e-3) / 1e12 / peak,
                "step_frac_of_burst_peak": step_flops / (ms_step * 1e-3) / 1e12 / peak_burst,
                "step_hbm_gbs": (8.0 * pairs + 4.0 * cfg["N"] * (cfg["D"] + cfg["d_img"])) / (ms_step * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks["hbm_gbs"]}

    # ------------------------------------------------------------------ end to end: host buffers through the public API
    from bayesvlm_b200.hostmem import pin, pinned_empty  # pinned staging buffers on the GPU's NUMA node

    img_host = EncoderResult(pin(t["img_e"], dev), pin(t["img_a"], dev))
    txt_host = EncoderResult(pin(t["txt_e"], dev), pin(t["txt_a"], dev))
    e2e_steps = max(1, min(K, 10))
    host_out = (pinned_empty((cfg["N"], cfg["C"]), device=dev), pinned_empty((cfg["N"], cfg["C"]), device=dev))
    for _ in range(2):
        model.predict_host(img_host, txt_host, batch_size=args.e2e_batch, out=host_out)
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = model.predict_host(img_host, txt_host, batch_size=args.e2e_batch, out=host_out)
    barrier_sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
    h2d = 4 * (t["img_e"].numel() + t["img_a"].numel() + t["txt_e"].numel() + t["txt_a"].numel())
    d2h = 4 * (res.mean.numel() + res.var.numel())
    e2e_err = float((res.mean[rows] - torch.from_numpy(ref_m).float()).abs().max())
    e2e = {"value": world * pairs / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "steps": e2e_steps, "max_abs_mean_err_384_rows": e2e_err,
           "api": "CLIP.predict_host (make_predictions data flow): pinned host -> device, kernels, device -> pinned host; "
                  "three-stream pipeline over image batches, caller-owned pinned result buffers",
           "note": "PCIe-bound: 764.5 MB cross the link per step; at N > 1 the ranks share the host's memory / PCIe root bandwidth"}
    del img_host, txt_host, res, out, img, txt, host_out
    state.clear()
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ KFAC estimation
    def kfac_leg(kc, cb_per_rank, likelihood, both_modalities, steps, total_cb=None, check=False):
        siglip = likelihood == "siglip"
        n_total = (total_cb if total_cb is not None else cb_per_rank * world) * kc["num_classes"]
        n_local = n_total // world
        data = kfac_inputs(kc, n_total, kc["seed"], device=dev, with_txt_act=both_modalities)  # every rank builds the same set
        e_img, e_txt, a_img = data[:3]
        a_txt = data[3] if both_modalities else None
        vlm = (SIGLIP(logit_scale=kc["logit_scale"], logit_bias=kc["logit_bias"], device=dev) if siglip
               else CLIP(logit_scale=LS, device=dev))
        kw = dict(num_classes=kc["num_classes"], batch_size=kc["batch_size"], device=str(dev), likelihood=likelihood,
                  distributed=use_dist)
        res_ = {}

        def one():
            res_["img"] = kfac_ggn(vlm, source_embeds=e_img, source_activations=a_img, target_embeds=e_txt, **kw)
            if both_modalities:
                res_["txt"] = kfac_ggn(vlm, source_embeds=e_txt, source_activations=a_txt, target_embeds=e_img, **kw)

        one()
        l1 = _lib.launch_count()
        ms = timed(one, steps, settle=1)
        n_launch = (_lib.launch_count() - l1) // (steps + 1)
        # per-kernel table from a separate pass (event pairs stay outside the headline region)
        _lib.timing_enable(True)
        one()
        barrier_sync()
        _lib.timing_enable(False)
        kk = _lib.timing_collect()
        C_, D_ = kc["num_classes"], kc["D"]
        d_i = kc["d_img"] + (1 if siglip else 0)
        d_t = kc["d_txt"] + (1 if siglip else 0)
        ggn = (4.0 * C_ * D_ + 2.0 * D_ * (D_ + 1) + 2.0 * D_ * D_) if siglip else (6.0 * C_ * D_ + 3.0 * D_ * (D_ + 1) + 2.0 * D_ * D_)
        per_sample = ggn + d_i * (d_i + 1.0)
        if both_modalities:
            per_sample += ggn + d_t * (d_t + 1.0)
        tf = per_sample * n_local / (ms * 1e-3) / 1e12
        leg = {"metric": "kfac_samples_per_s", "value": n_total / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "steps": steps,
               "workload": kc["name"], "class_batches_total": n_total // kc["num_classes"],
               "class_batches_per_gpu": n_local // kc["num_classes"], "modalities": "img + txt" if both_modalities else "img",
               "flop_per_sample": per_sample, "algo_tflops_per_gpu": tf, "frac_of_bf16_sustained": tf / peak,
               "frac_of_bf16_burst": tf / peak_burst,
               "kernels": {k: {"launches": v[0], "avg_ms": v[1] / v[0]} for k, v in kk.items()}, "gpu_launches": n_launch}
        if "syrk" in kk:  # K1 as its own roofline entry, symmetric flop count, live CUDA-event time per launch
            rows_per_launch = n_local * (2 if both_modalities else 1) / kk["syrk"][0]
            d_eff = d_i if not both_modalities else math.sqrt((d_i * (d_i + 1.0) + d_t * (d_t + 1.0)) / 2.0)
            sy = rows_per_launch * d_eff * (d_eff + 1.0) / (kk["syrk"][1] / kk["syrk"][0] * 1e-3) / 1e12
            leg["syrk_roofline"] = {"bound": "tensor", "achieved": sy, "peak": peak, "unit": "TFLOP/s", "frac": sy / peak,
                                    "rows_per_launch": rows_per_launch, "avg_launch_ms": kk["syrk"][1] / kk["syrk"][0],
                                    "flops": "n d (d+1), symmetric count"}
        A, B = res_["img"]
        assert torch.isfinite(A).all() and torch.isfinite(B).all()
        if check and world > 1:
            # the N-rank result equals the 1-rank result -- rank 0 recomputes every class batch alone
            ok = 1.0
            info = {}
            if rank == 0:
                A1, B1 = kfac_ggn(vlm, source_embeds=e_img, source_activations=a_img, target_embeds=e_txt,
                                  **{**kw, "distributed": False})
                ra = float((A - A1).norm() / A1.norm())
                rb = float((B - B1).norm() / B1.norm())
                # B: every class batch runs the identical kernel sequence on either side -> only the order of the fp32 sums
                # differs (1e-5).  A: the rank-local SYRK is ONE launch over the rank's rows, so the K extent of the tensor-core
                # accumulation (and the split-K partition) differs between the 1-rank and the N-rank run: both are within the
                # SYRK's 2e-4 of fp64 (tests/test_gpu_kfac.py::test_syrk_vs_fp64), measured difference 3e-5 -> 1e-4.
                info = {"rel_err_A": ra, "rel_err_B": rb, "tolerance_A": 1e-4, "tolerance_B": 1e-5,
                        "what": f"{world}-rank all-reduced factors vs the same {n_total // kc['num_classes']} class batches on rank 0 alone"}
                ok = 1.0 if (ra <= 1e-4 and rb <= 1e-5) else 0.0
            okt = torch.tensor([ok], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            leg["allreduce_check"] = info
            if float(okt.item()) < 0.5:
                raise RuntimeError(f"kfac all-reduce equivalence check failed: {info}")
        return leg, (e_img, e_txt, a_img, vlm, kw)

    ksteps = max(1, min(K, 5))
    kfac, kctx = kfac_leg(KFAC, args.kfac_class_batches, "info_nce", False, ksteps, check=True)
    kfac["scaling"] = "weak"
    kfac["factors"] = "A_img (768^2) + B_img (512^2), InfoNCE GGN + SYRK" + (", one all-reduce of [A||B]" if world > 1 else "")
    # the same through the reference's calling convention: HOST tensors in (pinned), B back on the CPU; kfac_ggn stages
    # every class batch one ahead on a copy stream
    e_img, e_txt, a_img, vlm, kw = kctx
    n_local = args.kfac_class_batches * KFAC["num_classes"]
    try:
        h_img, h_act, h_txt = (pin(x.cpu(), dev) for x in (e_img, a_img, e_txt))
        pinned_ok = 1.0
    except RuntimeError:  # not enough page-locked memory on a shared host: skip the (informational) leg on ALL ranks
        h_img = h_act = h_txt = None
        pinned_ok = 0.0
    del e_img, e_txt, a_img, kctx
    torch.cuda.empty_cache()
    if -max_over_ranks(-pinned_ok) > 0.5:
        kfac_ggn(vlm, source_embeds=h_img, source_activations=h_act, target_embeds=h_txt, **kw)
        barrier_sync()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            A, B = kfac_ggn(vlm, source_embeds=h_img, source_activations=h_act, target_embeds=h_txt, **kw)
        barrier_sync()
        kfac_e2e_ms = max_over_ranks((time.perf_counter() - t0) / ksteps * 1e3)
        kfac["e2e"] = {"value": n_local * world / (kfac_e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": kfac_e2e_ms,
                       "h2d_bytes_per_step": 4 * n_local * (2 * KFAC["D"] + KFAC["d_img"]), "d2h_bytes_per_step": 4 * KFAC["D"] ** 2,
                       "api": "kfac_ggn on pinned host tensors (the reference's calling convention), B returned on the CPU"}
    else:
        kfac["e2e"] = {"unavailable": "could not page-lock the host copies of the inputs on every rank"}
    del h_img, h_act, h_txt
    torch.cuda.empty_cache()

    # K1 alone (nothing else on the GPU): the rank's config-2 block of activations through ONE syrk_accumulate call. The
    # `syrk` entry of the KFAC legs above is timed while the SYRK shares the SMs with GGN pass 1 (side stream); this is the
    # kernel by itself -- tensor-core launch and whole operator (absmax + fp16 conversion + GEMM + mirror) separately.
    def syrk_leg(rows, d):
        from bayesvlm_b200.hessians import syrk_accumulate

        gs = torch.Generator(device=dev).manual_seed(KFAC["seed"] + 7 + rank)
        X = torch.randn(rows, d, generator=gs, device=dev)
        out_ = torch.zeros(d, d, device=dev)
        fn = lambda: syrk_accumulate(X, out=out_)
        fn()
        ms_op = timed(fn, 5, settle=1)
        _lib.timing_enable(True)
        fn()
        barrier_sync()
        _lib.timing_enable(False)
        kk = _lib.timing_collect()
        ms_gemm = kk["syrk"][1] / kk["syrk"][0]
        flop = float(rows) * d * (d + 1.0)
        ref = (X[:4096].double().T @ X[:4096].double())
        got = syrk_accumulate(X[:4096]).double()
        return {"rows": rows, "d": d, "flops": "n d (d+1), symmetric count", "unit": "TFLOP/s", "peak": peak,
                "gemm_launch_ms": ms_gemm, "gemm_tflops": flop / (ms_gemm * 1e-3) / 1e12, "frac": flop / (ms_gemm * 1e-3) / 1e12 / peak,
                "operator_ms": ms_op, "operator_tflops": flop / (ms_op * 1e-3) / 1e12,
                "operator_frac": flop / (ms_op * 1e-3) / 1e12 / peak,
                "operator_hbm_bytes": rows * d * (4 + 4 + 2 + 2), "rel_err_4096_rows_vs_fp64": float((got - ref).norm() / ref.norm())}

    try:
        kfac["syrk_solo"] = syrk_leg(KFAC["total_class_batches"] * KFAC["num_classes"] // world // (4 if args.quick else 1), KFAC["d_img"])
    except Exception as exc:
        kfac["syrk_solo"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()

    extra_legs = {}
    if not args.quick:
        for name, fn in (
                ("kfac_cfg2", lambda: kfac_leg(KFAC, None, "info_nce", True, max(1, min(K, 2)), total_cb=KFAC["total_class_batches"])[0]),
                ("kfac_h14", lambda: kfac_leg(KFAC_H14, 2, "info_nce", False, max(1, min(K, 3)))[0]),
                ("kfac_siglip", lambda: kfac_leg(KFAC_SIGLIP, 2, "siglip", False, max(1, min(K, 3)))[0])):
            try:
                extra_legs[name] = fn()
                extra_legs[name]["scaling"] = "strong" if name == "kfac_cfg2" else "weak"
            except Exception as exc:  # a secondary leg never takes the headline line down
                extra_legs[name] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()

    # ------------------------------------------------------------------ EPIG scoring (config 5; pool rows sharded, no collective)
    from bayesvlm_b200.epig import epig_from_logits_using_matmul, epig_from_probs_using_matmul
    from bayesvlm_b200.vlm import ProbabilisticLogits

    def epig_leg(cl, pool_total):
        ec = EPIG
        pool = max(1, pool_total // world)  # rows of this rank's shard (the last pool chunk may be ragged, as in the reference)
        geng = torch.Generator(device=dev).manual_seed(ec["seed"] + rank + 100 * cl)
        mk = lambda n: ProbabilisticLogits(torch.randn(n, cl, generator=geng, device=dev) * 2,
                                           torch.rand(n, cl, generator=geng, device=dev) * 3 + 0.1)
        lp, lt = mk(pool), mk(ec["target"])
        run = lambda: epig_from_logits_using_matmul(lp, lt, seed=0, num_samples=ec["K"], chunk_size=ec["chunk"])
        run()
        ms = timed(run, 1, settle=0)
        _lib.timing_enable(True)
        scores = run()
        barrier_sync()
        _lib.timing_enable(False)
        ek = _lib.timing_collect()
        assert torch.isfinite(scores).all()
        epairs = float(pool) * ec["target"]
        n_chunks = math.ceil(pool / ec["chunk"])
        kp = int(_lib.lib.bvlm_epig_operand_k(ec["K"]))
        # algorithmic bytes of the fused sample / permute / marginal-entropy pass: noise + logits read, operand + entropies written
        rows_all = n_chunks * ec["target"] + pool
        prep_bytes = rows_all * (ec["K"] * cl * 4 + 2 * cl * 4 + cl * kp * 2 + 2)
        prep_ms = ek.get("epig_prepare", (0, 0.0))[1]
        joint_ms = ek.get("epig_joint", (0, 0.0))[1]
        leg = {"metric": "epig_pool_target_pairs_per_s", "value": world * epairs / (ms * 1e-3), "unit": "pairs/s", "ms": ms,
               "workload": ec["name"], "classes": cl, "pool_rows_per_gpu": pool, "pool_rows_total": pool * world,
               "target_rows": ec["target"], "joint_kernel_ms": joint_ms,
               "joint_log_evals_per_s": epairs * cl ** 2 / max(joint_ms * 1e-3, 1e-12),
               "joint_frac_of_mufu_roof": epairs * cl ** 2 / max(joint_ms * 1e-3, 1e-12) / (16.0 * 148 * 1.965e9),
               "reductions": {"what": "E0 sample -> fp16 -> permuted GEMM operand -> E1 marginal entropy; target set + pool chunk in ONE launch per chunk",
                              "launches": ek.get("epig_prepare", (0, 0.0))[0], "bytes": prep_bytes, "ms": prep_ms,
                              "gbs": prep_bytes / max(prep_ms * 1e-3, 1e-12) / 1e9,
                              "frac": prep_bytes / max(prep_ms * 1e-3, 1e-12) / 1e9 / peaks["hbm_gbs"], "peak_gbs": peaks["hbm_gbs"]}}
        return leg, (lp, lt)

    epig, (lp, lt) = epig_leg(10, EPIG["pool"])
    # parity of the scored path: one pool chunk against the reference's operation sequence (torch CUDA kernels) on this GPU
    try:
        from oracle import torch_port as T

        n_chk = min(EPIG["chunk"], lp.mean.shape[0])
        s_ours = epig_from_logits_using_matmul(ProbabilisticLogits(lp.mean[:n_chk], lp.var[:n_chk]), lt, seed=0,
                                               num_samples=EPIG["K"], chunk_size=EPIG["chunk"])
        s_ref = T.epig_from_logits(lp.mean[:n_chk], lp.var[:n_chk], lt.mean, lt.var, seed=0, num_samples=EPIG["K"],
                                   chunk_size=EPIG["chunk"])
        d = (s_ours - s_ref).abs()
        top_a = set(torch.argsort(s_ours, descending=True)[:50].tolist())
        top_b = set(torch.argsort(s_ref, descending=True)[:50].tolist())
        epig["parity"] = {"rows": int(n_chk), "exact_match_rate": float((d == 0).float().mean()), "max_abs_diff": float(d.max()),
                          "top50_identical": top_a == top_b, "top50_symmetric_difference": len(top_a ^ top_b),
                          "against": "oracle/torch_port.epig_from_logits (the reference's sequence, torch CUDA kernels) on the same GPU, shared device RNG"}
        if epig["parity"]["exact_match_rate"] < 0.95:
            raise RuntimeError(f"EPIG parity check failed: {epig['parity']}")
    except RuntimeError:
        raise
    except Exception as exc:
        epig["parity"] = {"error": repr(exc)[:300]}
    del lp, lt
    torch.cuda.empty_cache()
    if not args.quick:
        try:
            epig["cl65"], _ = epig_leg(65, EPIG["pool_cl65"])
        except Exception as exc:
            epig["cl65"] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()
    clocks.__exit__()

    # ------------------------------------------------------------------ reference algorithm on the host cores (rank 0, N=1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, dt = cpu_predictive_rate(cfg, cfg["N"], steps=5, warmup=1)
        krate, kdt = cpu_kfac_rate(KFAC, data_batches=8)
        cpu_baseline = {"value": rate, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"5 full {cfg['N']} x {cfg['C']} calls of the reference algorithm "
                                  f"(oracle/torch_port.predictive, torch CPU fp32, {dt * 1e3:.0f} ms/call)",
                        "kfac": {"value": krate, "unit": "samples/s",
                                 "sample": f"reference double loop, 1 class batch of 32768 targets x 8 data batches of 5 ({kdt:.1f} s)"}}
        kfac["vs_cpu_port"] = kfac["value"] / krate
        try:
            cpu_baseline["same_gpu_torch_eager"] = gpu_eager_rates(cfg, KFAC, dev)
        except Exception as exc:  # informational leg: never fail the bench line on it
            cpu_baseline["same_gpu_torch_eager"] = {"unavailable": repr(exc)[:200]}
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": "predictive_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": PREC_DTYPE.get(model.precision, model.precision),
            "data": "synthetic",
            "config": pred_config(cfg, world, {"outputs": "logit mean + variance fp32",
                                               "precision": PREC_TEXT.get(model.precision, model.precision),
                                               "l2": "per-step inputs 359 MB + outputs 400 MB exceed the 126 MB L2 (no flush needed)",
                                               "sharding": f"images row-sharded over {world} rank(s), classes replicated, no collective"}),
            "roofline": roofline, "parity_check": parity, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks.summary(), "with_probs": with_probs, "kfac": kfac, "epig": epig,
        }
        line.update(extra_legs)
        print(json.dumps(line), flush=True)


PREC_DTYPE = {"fp16x3": "f16x3->f32", "fp16+fp8": "f16+e4m3->f32", "fp16": "f16->f32"}
PREC_TEXT = {
    "fp16x3": "fp16 hi/lo split mean GEMM (3 tensor-core passes, fp32 accumulate); fp16 quadratic form",
    "fp16+fp8": "fp16 mean GEMM + E4M3 error-compensation K phase into the same fp32 accumulator (max abs logit error 0.2 of "
                "the 1e-3 tolerance); fp16 quadratic form",
    "fp16": "single fp16 mean GEMM (fp32 accumulate); fp16 quadratic form",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-batch", type=int, default=2048,
                    help="image batch of the three-stream host pipeline (scripts/e2e_sweep.py: 2048 is the measured optimum)")
    ap.add_argument("--kfac-class-batches", type=int, default=2, help="class batches of 32768 per GPU per KFAC step (weak-scaling leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the secondary legs (config 2 in full, H-14, SigLIP, EPIG Cl=65)")
    ap.add_argument("--precision", default=None, help="mean-logit GEMM precision: fp16x3 | fp16+fp8 | fp16 (default: the library's)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
