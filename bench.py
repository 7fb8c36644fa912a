#!/usr/bin/env python
"""bench.py — nucleotides/sec fwd+bwd of HyenaDNA @ 1 M bp on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (torchrun for N > 1)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU oracle of the reference path

One "step" = one pass of the hot path over one batch of synthetic nucleotides: tokenise raw bytes
(hy_tokenize) -> 8-layer d_model=256 HyenaDNA forward (our fused HyenaOperator kernels inside a plain
PyTorch backbone) -> next-token cross-entropy -> backward -> DP gradient all-reduce (N > 1) -> AdamW.
`value`  : inputs (raw bytes) resident in HBM, loss kept on device; EXACTLY K steps after W >= 3 warm-up steps, CUDA events,
           barrier + synchronize on both sides, max over ranks, no per-kernel instrumentation.
`roofline`: the same K steps once more with CUDA-event brackets around every kernel of ours (and every exchange at N > 1):
           launch durations, breakdown, exchange time; `roofline.instrumented_ms_per_step` is that pass's step time.
`e2e`    : the same step through the public API with HOST (pinned) buffers: H2D copy of the bytes and a
           D2H read of the loss inside the timed region, every step.
N > 1, `--partition channels` (the default for the single-sequence 1 M workload, BASELINE.json configs[3]): ONE 1 M-nt
sequence split over the ranks — sequence chunks outside the operator core, channel slabs inside it, two all-to-alls
per layer and direction (dna_b200.dp.ChannelPartition) — strong scaling; the line also carries the batch-sharded
weak-scaling number of the same run under `batch_dp`.  `--partition batch`: every rank processes its own sequence
(no collective inside the operator), one flat fp32 gradient all-reduce per step, weak scaling.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric is quoted on (fits one GPU)
    "hyenadna-large-1m": dict(n_layer=8, d_model=256, d_inner=1024, seqlen=1_000_000, batch=1),
    # BASELINE.json configs[2] / [1] / [0] (parity-test sizes; selectable for experiments)
    "hyenadna-medium-160k": dict(n_layer=8, d_model=256, d_inner=1024, seqlen=160_000, batch=1),
    "hyenadna-small-32k": dict(n_layer=4, d_model=256, d_inner=1024, seqlen=32_768, batch=8),
    "hyenadna-tiny-1k": dict(n_layer=2, d_model=128, d_inner=512, seqlen=1024, batch=8),
}
LAYER_CFG = dict(emb_dim=5, filter_order=64, short_filter_order=3, modulate=True, w=10, lr=6e-4, wd=0.0, lr_pos_emb=0.0)
METRIC = "nucleotides/sec fwd+bwd HyenaDNA @1M bp"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hyenadna-large-1m", choices=list(WORKLOADS))
    ap.add_argument("--seqlen", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--checkpoint", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--partition", default="auto", choices=["auto", "batch", "channels"],
                    help="N > 1: 'channels' = one sequence split over the ranks (strong scaling), 'batch' = one sequence "
                         "per rank (weak scaling); auto = channels for the single-sequence 1 M workload, else batch")
    ap.add_argument("--cpu-sample-len", type=int, default=65536)
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N > 1, channels: skip the batch_dp leg")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="capture the whole step (tokenise .. AdamW) in ONE CUDA graph and replay it: the small workloads "
                         "are host-bound otherwise (252 launches of ours + ATen per step); auto = on for hyenadna-tiny-1k at N = 1")
    return ap.parse_args()


def synth_bytes(B, L, seed):
    """uniform random A/C/G/T with 1 % N, as raw ASCII."""
    import numpy as np
    rng = np.random.default_rng(seed)
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(B, L))].copy()
    arr[rng.random((B, L)) < 0.01] = ord("N")
    return arr


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the reference) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, sample_len, steps, warmup):
    """The oracle (oracle/: CPU restatement of standalone_hyenadna, fp32) on all host cores, on a bounded sample of the
    workload: B = 1 sequence of min(sample_len, seqlen) nucleotides, tokenise + forward + loss + backward."""
    import torch
    from oracle import hyena_model_oracle as MO
    from oracle import hyena_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L = min(sample_len, cfg["seqlen"])
    B = 1
    sd0 = MO.init_state_dict(cfg["d_model"], cfg["n_layer"], cfg["d_inner"], 16, L + 2, emb_dim=LAYER_CFG["emb_dim"],
                             filter_order=LAYER_CFG["filter_order"], w=LAYER_CFG["w"], seed=0)
    buffers = ("pos_emb.z", "pos_emb.t", "modulation.deltas")
    sd = {k: v.requires_grad_(not k.endswith(buffers)) for k, v in sd0.items()}
    text = synth_bytes(B, L, 0)
    times = []
    for it in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        ids = torch.tensor([O.tokenize_ref(bytes(row).decode(), L + 1) for row in text])
        loss = MO.lm_loss(ids[:, :-1], ids[:, 1:], sd, n_layer=cfg["n_layer"], l_max=L + 2, shift=0.05)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    best = min(times)
    mean = sum(times) / len(times)
    sample = (f"oracle (CPU torch restatement of standalone_hyenadna) model {cfg['n_layer']}L d{cfg['d_model']} fp32, "
              f"tokenise+fwd+bwd of B={B} x L={L} nt (bounded sample of the {cfg['seqlen']}-nt workload; the FFT is "
              f"O(L log L) and a 1 M-nt step no longer fits the caches, so nt/s at this length OVERSTATES the CPU at 1 M), "
              f"{warmup} warm-up + {len(times)} timed steps, mean (best {B * L / best:.0f} nt/s)")
    return dict(value=B * L / mean, best=B * L / best, ms_per_step=mean * 1e3, cores=cores, sample=sample, L=L, B=B)


def gpu_reference_run(cfg, dev, steps=3, warmup=1):
    """GPU baseline: the reference path itself — the oracle's eager torch restatement, i.e.
    torch.fft (cuFFT) long convolution, cuDNN short filter, ATen elementwise gates, under bf16 autocast like the
    reference's 16-bit training — on the same B200, at the largest power-of-two fraction of the workload that fits."""
    import torch
    from oracle import hyena_model_oracle as MO
    L = cfg["seqlen"]
    last_err = None
    while L >= 16384:
        try:
            sd0 = MO.init_state_dict(cfg["d_model"], cfg["n_layer"], cfg["d_inner"], 16, L + 2, emb_dim=LAYER_CFG["emb_dim"],
                                     filter_order=LAYER_CFG["filter_order"], w=LAYER_CFG["w"], seed=0)
            buffers = ("pos_emb.z", "pos_emb.t", "modulation.deltas")
            sd = {k: v.to(dev).requires_grad_(not k.endswith(buffers)) for k, v in sd0.items()}
            ids = torch.randint(7, 11, (1, L + 1), device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for it in range(warmup + steps):
                if it == warmup:
                    torch.cuda.synchronize()
                    ev[0].record()
                for v in sd.values():
                    v.grad = None
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = MO.lm_loss(ids[:, :-1], ids[:, 1:], sd, n_layer=cfg["n_layer"], l_max=L + 2, shift=0.05)
                loss.backward()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / steps
            peak = torch.cuda.max_memory_allocated() / 2 ** 30
            del sd, loss
            torch.cuda.empty_cache()
            return {"value": L / (ms * 1e-3), "unit": "nt/s", "ms_per_step": ms, "seqlen": L, "peak_mem_gib": round(peak, 1),
                    "kind": "oracle restatement of the reference's eager path (torch.fft / cuFFT, cuDNN conv1d, ATen gates) on "
                            "the same GPU, bf16 autocast, fwd + loss + bwd, no optimizer; largest seqlen (workload halved "
                            "until it fits) — same model shape",
                    "steps": steps}
        except torch.cuda.OutOfMemoryError as e:      # noqa: PERF203
            last_err = str(e)[:80]
            sd = loss = None
            torch.cuda.empty_cache()
            L //= 2
    return {"unavailable": last_err or "does not fit"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    r = cpu_reference_run(cfg, args.cpu_sample_len, steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "nt/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "n_layer": cfg["n_layer"], "d_model": cfg["d_model"], "seqlen": cfg["seqlen"],
                   "batch_per_gpu": cfg["batch"], "note": "CPU run on a bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {"value": r["value"], "unit": "nt/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "nt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, without perturbing it.

    Polling NVML / nvidia-smi concurrently with the launches stalls them for 100s of ms per query on this
    driver (measured: 458 -> 926 ms/step at 4 Hz, 650 ms/step with `nvidia-smi -lms 100`).  So:
      * the SM clock is measured IN-STREAM once per step by `hy_clock_probe` (clock64 vs globaltimer);
      * NVML (clock, max clock, throttle-reason mask) is polled a few times right after the last timed
        step has been ENQUEUED, while the GPU is still draining the queue — under load, but with no launch
        left to delay."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, device):
        import torch
        self.index = index
        self.nvml_clocks, self.max_clock, self.reasons = [], None, set()
        self.probe_buf = torch.zeros((512, 2), dtype=torch.int64, device=device)
        self.n_probe = 0
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_clock = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def probe(self):
        """enqueue one in-stream SM clock measurement (call once per step)"""
        from dna_b200 import kernels as K
        if self.n_probe < self.probe_buf.shape[0]:
            K.clock_probe(self.probe_buf[self.n_probe])
            self.n_probe += 1

    def poll_nvml(self, times=3, gap=0.03):
        if self.h is None:
            return
        for _ in range(times):
            try:
                self.nvml_clocks.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(gap)

    def result(self):
        mhz = []
        if self.n_probe:
            for cyc, ns in self.probe_buf[:self.n_probe].cpu().tolist():
                if ns > 0 and cyc > 0:
                    mhz.append(cyc / ns * 1e3)
        out = {"sm_mhz": round(statistics.median(mhz), 1) if mhz else (statistics.median(self.nvml_clocks) if self.nvml_clocks else None),
               "sm_max_mhz": self.max_clock, "reasons": sorted(self.reasons),
               "method": "per-step in-stream clock64/globaltimer probe (median); NVML clock+reasons polled while the last timed steps drain",
               "samples": len(mhz), "nvml_sm_mhz_under_load": self.nvml_clocks}
        return out


def run_ours(args, cfg):
    import gc
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from dna_b200 import kernels as K
    from dna_b200.dp import ChannelPartition, FlatGradAllReduce, set_channel_partition
    from dna_b200.standalone import HyenaDNAModel
    from dna_b200.tokenizer import CharacterTokenizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: hyena-b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L = cfg["batch"], cfg["seqlen"]
    D = cfg["d_model"]
    bf16 = args.dtype == "bf16"
    partition = args.partition
    if partition == "auto":
        partition = "channels" if (world > 1 and B == 1 and args.workload == "hyenadna-large-1m" and L % world == 0
                                   and D % world == 0) else "batch"
    if world == 1:
        partition = "batch"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(partition, with_e2e=True, sampler_on=True):
        """warm-up + timed device-resident steps (+ timed end-to-end steps) of one partition mode"""
        channels = partition == "channels"
        torch.manual_seed(2222)

        def make_model(ckpt):
            m = HyenaDNAModel(d_model=D, n_layer=cfg["n_layer"], d_inner=cfg["d_inner"], vocab_size=12,
                              pad_vocab_size_multiple=8, embed_dropout=0.1, resid_dropout=0.0, lm_head=True, checkpoint_blocks=ckpt,
                              layer=dict(l_max=L + 2, **LAYER_CFG)).to(dev)
            m.train()
            return m

        # rough activation footprint (bytes) without checkpointing: ~60 B/nt/channel... measured in DESIGN.md
        est = cfg["n_layer"] * B * L * D * (64 if bf16 else 110) / (world if channels else 1)
        ckpt = args.checkpoint == "on" or (args.checkpoint == "auto" and est > 140e9)
        part = ChannelPartition() if channels else None
        state = {}

        def build(ckpt):
            state["model"] = make_model(ckpt)
            if channels:
                set_channel_partition(state["model"], part)
            state["reducer"] = FlatGradAllReduce(state["model"].parameters())
            state["opt"] = torch.optim.AdamW(state["model"].parameters(), lr=6e-4, weight_decay=0.1, fused=True)

        build(ckpt)
        n_params = sum(p.numel() for p in state["model"].parameters())
        if channels:
            # one sequence for the whole job: this rank tokenises positions [lo, hi] (one byte of overlap: the target of
            # its last position); the last rank's final target is the trailing [SEP]
            lo, hi = part.chunk(L)
            Lc = hi - lo
            last = rank == world - 1
            host_np = synth_bytes(B, L, seed=0)[:, lo:(hi if last else hi + 1)].copy()
            tok = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=Lc + 1)
            denom = float(B * L)
        else:
            Lc, last = L, True
            host_np = synth_bytes(B, L, seed=rank)
            tok = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=L + 1)
            denom = float(B * L)
        host = torch.from_numpy(host_np).pin_memory()
        dev_bytes = host.to(dev)
        h2d_bytes = host.numel()

        def step(src_bytes, read_loss):
            ids = tok.encode_bytes_cuda(src_bytes, None, Lc + 1, add_special_tokens=last)     # [B, Lc+1]
            data, target = ids[:, :-1], ids[:, 1:]
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
                logits = state["model"](data)
            loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), target.reshape(-1), reduction="sum") / denom
            state["reducer"].zero()
            loss.backward()
            # channels: every rank holds the gradient contribution of its chunk / slab of the ONE sequence -> sum;
            # batch: mean over the ranks' sequences
            state["reducer"].allreduce(average=not channels)
            state["opt"].step()
            if read_loss:
                return float(loss.item())          # D2H read of the step's result
            return loss

        use_graph = world == 1 and (args.graph == "on" or (args.graph == "auto" and args.workload == "hyenadna-tiny-1k"))
        gstate = {}

        def device_step():
            if "graph" in gstate:
                gstate["graph"].replay()
                return gstate["loss"]
            return step(dev_bytes, False)

        def e2e_step():
            if "graph" in gstate:
                dev_bytes.copy_(host, non_blocking=True)     # H2D into the graph's static input buffer
                gstate["graph"].replay()
                return float(gstate["loss"].item())          # D2H read of the step's result
            return step(host.to(dev, non_blocking=True), True)

        def timed(fn, n, sampler=None):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            walls = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                if sampler is not None:
                    sampler.probe()
                walls.append(round((time.perf_counter() - t0) * 1e3, 1))
            e1.record()
            if sampler is not None:
                sampler.poll_nvml()          # GPU still draining the last steps: under load, nothing left to stall
            barrier()
            if os.environ.get("HY_BENCH_DEBUG"):
                st = torch.cuda.memory_stats()
                print(f"[bench debug] rank {rank} {partition} {fn.__name__} host ms per step: {walls} | cudaMalloc calls "
                      f"{st.get('num_device_alloc')}, cudaFree calls {st.get('num_device_free')}, retries {st.get('num_alloc_retries')}, "
                      f"reserved {st.get('reserved_bytes.all.current', 0) / 2**30:.1f} GiB", file=sys.stderr, flush=True)
            ms = e0.elapsed_time(e1) / n
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        try:
            for _ in range(max(args.warmup, 3)):
                device_step()
            torch.cuda.synchronize()
        except torch.cuda.OutOfMemoryError:
            if ckpt:
                raise
            state.clear()
            torch.cuda.empty_cache()
            ckpt = True
            build(True)
            for _ in range(max(args.warmup, 3)):
                device_step()
            torch.cuda.synchronize()

        if use_graph:
            # whole-step CUDA graph: static input bytes, gradients live in the flat buffer (static), capturable AdamW
            state["opt"] = torch.optim.AdamW(state["model"].parameters(), lr=6e-4, weight_decay=0.1, fused=True, capturable=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(dev_bytes, False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # per-kernel CUDA-event timings (the roofline's launch duration) come from an eager pass: replays do not
            # run the host-side event brackets
            K.enable_timing(True)
            K.drain_timing()
            for _ in range(args.steps):
                step(dev_bytes, False)
            gstate["kt"] = K.drain_timing()
            K.enable_timing(False)
            graph = torch.cuda.CUDAGraph()
            n0 = K.launch_count()
            with torch.cuda.graph(graph):
                gstate["loss"] = step(dev_bytes, False)
            gstate["launches"] = K.launch_count() - n0      # kernels of ours inside ONE replay
            gstate["graph"] = graph
            for _ in range(3):
                device_step()
            torch.cuda.synchronize()

        # the cyclic GC ran mid-step (hundreds of ms with GB-sized graphs alive): collect now, keep it off while timing
        gc.collect()
        gc.disable()
        sampler = ClockSampler(local, dev) if (rank == 0 and sampler_on) else None
        launches0 = K.launch_count()
        # pass 1 — the reported step time: no per-kernel event brackets (at N = 8 the channel partition enqueues ~1000
        # launches per 45 ms step; ~700 extra event records per step made this pass host-bound on a busy host: 72 ms
        # against 45.6 ms for the un-instrumented e2e pass of the same run, profiles/r02s_*)
        ms = timed(device_step, args.steps, sampler)
        launches = (K.launch_count() - launches0 - (args.steps if sampler is not None else 0)) // max(args.steps, 1)
        # pass 2 — the same K steps again with CUDA-event brackets around every kernel of ours and every exchange: the
        # roofline's launch durations, the breakdown and the exchange time come from here, `ms_instrumented` is its step time
        K.enable_timing(True)
        K.drain_timing()
        if part is not None:
            part.bytes_sent = 0
            part.enable_timing(True)
        ms_instr = timed(device_step, args.steps) if "graph" not in gstate else ms
        kt = K.drain_timing()
        K.enable_timing(False)
        if "graph" in gstate:
            kt = gstate["kt"]
        res = {"ms": ms, "ms_instr": ms_instr, "kt": kt, "ckpt": ckpt, "n_params": n_params, "h2d_bytes": int(h2d_bytes),
               "graph": "graph" in gstate, "launches": gstate["launches"] if "graph" in gstate else launches,
               "clocks": sampler.result() if sampler is not None else {}, "grad_bytes": state["reducer"].nbytes}
        if part is not None:
            part.check()
            res["a2a_ms"] = part.drain_timing() / max(args.steps, 1)
            res["a2a_bytes"] = part.bytes_sent // max(args.steps, 1)
            res["exchange_backend"] = part.backend
            part.enable_timing(False)
        if with_e2e:
            e2e_step()
            res["ms_e2e"] = timed(e2e_step, args.steps)
        gc.enable()
        res["peak_mem"] = torch.cuda.max_memory_allocated() / 2 ** 30
        state.clear()
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        return res

    r = measure(partition)
    ms, ms_e2e, kt = r["ms"], r["ms_e2e"], r["kt"]
    channels = partition == "channels"
    seqs = B if channels else world * B                   # sequences the whole job processes per step
    value = seqs * L / (ms * 1e-3)
    e2e_value = seqs * L / (ms_e2e * 1e-3)

    # roofline of the fused long-conv + gating kernel family (SURVEY §8d algorithmic bytes), per launch on ONE rank:
    # a rank's launch covers B x (D or D / world) channel rows of L positions
    s = 2 if bf16 else 4
    rows = B * (D // world if channels else D)
    alg_bytes_layer = 11 * s * rows * L + 12 * (rows // B) * L
    fam = ("spectrum", "conv_fwd", "conv_bwd", "conv_dk")
    fam_ms_step = sum(kt[t][1] for t in fam if t in kt) / max(args.steps, 1)
    per_call_ms = fam_ms_step / cfg["n_layer"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_layer / (per_call_ms * 1e-3) / 1e9 if per_call_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic_per_layer(rows, L, bf16)
    # compute view (at L >= 256 k report both): 6 packed-real transforms per row and layer
    # (filter spectrum, g forward, y inverse | dy forward, dg inverse, dk inverse), 5 M log2 M flops each, M = N / 2 complex points
    try:
        M = K.fft_len(L)          # complex transform length the library uses for L (a power of two, or 3 / 5 times one)
    except Exception:
        M = 1
        while M < L:
            M *= 2
    import math
    fft_flops = 6 * rows * 5 * M * math.log2(M) if rows else 0
    tflops = fft_flops / (per_call_ms * 1e-3) / 1e12 if per_call_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel": "fused long-conv + gating, one layer fwd+bwd = hy_filter_spectrum + hy_conv_fwd + hy_conv_bwd + hy_conv_dk",
                "algorithmic_bytes_per_launch": alg_bytes_layer, "ms_per_launch": per_call_ms,
                "flops": {"fft_tflops_achieved": tflops, "fp32_cuda_core_peak_tflops": 72.0,
                          "frac_of_fp32_peak": tflops / 72.0,
                          "note": "6 complex FFTs of M = %d points per channel row and layer (5 M log2 M flops each) on the CUDA "
                                  "cores; 72 TFLOP/s = 148 SMs x 128 FMA lanes x 2 x 1.9 GHz" % M},
                "share_of_step": fam_ms_step / r["ms_instr"] if r["ms_instr"] > 0 else None,
                "instrumented_ms_per_step": r["ms_instr"],
                "timing": "launch durations, breakdown and share: CUDA events around every kernel of ours during a second "
                          "pass of the same K steps (instrumented_ms_per_step); ms_per_step / value: the first, un-instrumented pass",
                "breakdown_ms_per_step": {t: kt[t][1] / max(args.steps, 1) for t in kt}}

    if channels:
        exch = ("our pull kernels over NVLink peer memory (TMA bulk copies, hy_exchange.cu)" if r.get("exchange_backend") == "peer"
                else "packed NCCL all_to_all_single (peer memory unavailable on this box)")
        par = (f"cp{world}: ONE sequence, sequence chunks of {L // world} nt outside the operator core, channel slabs of "
               f"{D // world} inside it; transposing exchanges between the two ({4 * cfg['n_layer']} per step) by {exch}; "
               f"flat NCCL grad all-reduce (sum)")
    else:
        par = f"dp{world} (batch-sharded, flat NCCL grad all-reduce)"
    line = {
        "metric": METRIC, "value": value, "unit": "nt/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if channels else "weak", "vs_baseline": None,
        "dtype": "bf16 activations / fp32 FFT" if bf16 else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "n_layer": cfg["n_layer"], "d_model": D, "d_inner": cfg["d_inner"], "seqlen": L,
                   "batch_per_gpu": (B / world if channels else B), "global_batch": seqs, "partition": partition, "parallelism": par,
                   "step": "tokenize + fwd (embed_dropout 0.1, resid_dropout 0 as hg38_hyena.yaml:12-13) + CE loss + bwd + "
                           "grad all-reduce + AdamW", "params": r["n_params"],
                   "activation_checkpointing": bool(r["ckpt"]), "cuda_graph": bool(r["graph"]), "l2": "inputs_larger_than_L2 (GBs of activations per step)",
                   "peak_mem_gib": round(r["peak_mem"], 1)},
        "clocks": r["clocks"],
        "e2e": {"value": e2e_value, "unit": "nt/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": r["h2d_bytes"] * (world if channels else 1),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(r["launches"]),
        "roofline": roofline,
    }
    if channels:
        line["collective"] = {"all_to_all_ms_per_step": r["a2a_ms"], "all_to_all_bytes_sent_per_rank_per_step": int(r["a2a_bytes"]),
                              "all_to_all_calls_per_step": 4 * cfg["n_layer"], "grad_allreduce_bytes": int(r["grad_bytes"]),
                              "backend": r.get("exchange_backend"),
                              "note": "device time between CUDA events around every exchange call of rank 0, on the stream "
                                      "it was issued on (includes waiting for the slowest rank); the filter-path exchanges "
                                      "run on a side stream beside compute, the others are not overlapped"}
        if not args.no_secondary:
            r2 = measure("batch", with_e2e=False, sampler_on=False)
            line["batch_dp"] = {"value": world * B * L / (r2["ms"] * 1e-3), "unit": "nt/s", "ms_per_step": r2["ms"], "scaling": "weak",
                                "global_batch": world * B, "peak_mem_gib": round(r2["peak_mem"], 1),
                                "parallelism": f"dp{world} (one sequence per rank, flat NCCL grad all-reduce)"}
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1:
        # tokenizer kernel against the HBM roofline (1 B read + 8 B written per nucleotide), on a batch
        # larger than L2 (the step's own 1 M-nt call is launch-latency bound and L2 resident)
        try:
            tokB, tokL = 256, 1 << 20     # long enough (~0.4 ms) that the ~0.1 ms of host work per call stays hidden
            tb = torch.randint(65, 85, (tokB, tokL), dtype=torch.uint8, device=dev)
            tokz = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=tokL + 1)
            for _ in range(3):
                ids_ = tokz.encode_bytes_cuda(tb, None, tokL + 1, add_special_tokens=True)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(10):
                ids_ = tokz.encode_bytes_cuda(tb, None, tokL + 1, add_special_tokens=True)
            ev1.record()
            torch.cuda.synchronize()
            tms = ev0.elapsed_time(ev1) / 10
            tgb = 9.0 * tokB * tokL / (tms * 1e-3) / 1e9
            line["tokenizer"] = {"achieved": tgb, "unit": "GB/s", "peak": peak, "frac": tgb / peak, "ms": tms,
                                 "workload": f"{tokB} x {tokL} nt ({tokB} MB in, {8 * tokB} MB of int64 ids out), 9 B per nucleotide"}
            del tb, ids_
        except Exception as e:  # noqa: BLE001
            line["tokenizer"] = {"unavailable": str(e)[:100]}
    if rank == 0:
        if world == 1 and not args.no_gpu_baseline:
            line["gpu_baseline"] = gpu_reference_run(cfg, dev)
        if world == 1 and not args.no_cpu_baseline:
            rc = cpu_reference_run(cfg, args.cpu_sample_len, steps=args.cpu_steps, warmup=1)
            line["cpu_baseline"] = {"value": rc["value"], "unit": "nt/s", "cores": rc["cores"], "kind": "port", "sample": rc["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


NCU_FAMILY_CSV = os.path.join(ROOT, "profiles", "r02q_ncu_full_longconv_family_1m_128rows.csv")


def ncu_traffic_per_layer(rows, L, bf16):
    """dram__bytes_read.sum + dram__bytes_write.sum of one layer's long-conv launch family, from the committed
    `ncu --set full` capture of the same kernels (128 rows, L = 1 000 000, bf16; every kernel of the family is
    linear in the row count, so the capture is scaled by rows / 128). None for other workloads."""
    if not (bf16 and L == 1_000_000 and os.path.exists(NCU_FAMILY_CSV)):
        return None, None
    import csv
    with open(NCU_FAMILY_CSV) as f:
        r = list(csv.reader(f))
    hdr, units, data = r[0], r[1], r[2:]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[ir]]
    tot = sum(float(x[ir]) + float(x[iw]) for x in data) * scale
    return tot * rows / 128.0, "profiles/" + os.path.basename(NCU_FAMILY_CSV) + " (ncu --set full, 128 rows) x rows/128"


def main():
    args = parse()
    cfg = dict(WORKLOADS[args.workload])
    if args.seqlen:
        cfg["seqlen"] = args.seqlen
    if args.batch:
        cfg["batch"] = args.batch
    if args.layers:
        cfg["n_layer"] = args.layers
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
