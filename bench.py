#!/usr/bin/env python
"""bench.py -- probit-ELBO fwd+bwd throughput (label-samples/s) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload eurlex] [--z Z] [--impl reference]

One "step" = one pass of the hot path over one batch: draw the (S,B,Z) noise (Philox, on device), the
forward of compute_loss (mpvae.py:145-210) and its full backward to logits / R / mu / logvar; at N > 1 every
rank owns B rows (weak scaling) and the step ends with the NCCL all-reduce of g_R (the path's one exchange).
`value` times that with the inputs resident in HBM; `e2e` times the same call through the public
`mpvae_b200.compute_loss` API starting from pinned HOST buffers (H2D of the step's inputs and D2H of the loss
inside the timed region).  Prints ONE JSON line (rank 0).

`--impl reference` times the reference's algorithm on the host cores instead (the oracle port of
/root/reference/mpvae.py -- the reference itself is Python and does not travel to the GPU box), on a bounded
row sample of the same workload.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "probit-ELBO fwd+bwd label-samples/s"
UNIT = "label-samples/s"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="eurlex", help="one of mpvae_b200.synth.SHAPES")
    ap.add_argument("--z", type=int, default=None, help="rank of R (default: the workload's, z_dim = label_dim)")
    ap.add_argument("--batch", type=int, default=None, help="rows per GPU (default: the workload's)")
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the full-training-step (steps/s) leg")
    ap.add_argument("--engine", default="auto", choices=["auto", "fma", "tensor"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nvls", "nccl"],
                    help="N>1: how g_R is summed over ranks (peer = inside the backward over NVLink peer memory, "
                         "mpvae_b200.peer.PeerRing, falling back to nccl if the ring cannot be set up on every rank; "
                         "nccl = all-reduce after the backward)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []          # (arrival time, csv line); only those inside [t0, t1] are reported
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else float("-inf")
        t1 = self.t1 if self.t1 is not None else float("inf")
        for when, ln in self.lines:
            if not (t0 <= when <= t1):
                continue          # nvidia-smi was started before the warm-up so that it is already looping here
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(a):
    from mpvae_b200 import synth
    sh = synth.SHAPES[a.workload]
    Z = a.z if a.z is not None else sh.z_dim
    B = a.batch if a.batch is not None else sh.batch
    return sh, sh.label_dim, Z, B, sh.n_sample


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(L, Z, S, rows, seed=0):
    """One fwd(+bwd) of the reference algorithm (oracle port, O(L^2) pairwise ranking loss, all host threads)
    on `rows` rows of the workload.  Returns seconds."""
    import torch
    from mpvae_b200 import synth
    from oracle import probit_elbo_oracle as orc
    inp = synth.loss_inputs(L, Z, rows, S, seed=seed + 1, label_rate=min(0.1, 20.0 / L))
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    noise = t.pop("noise")
    t0 = time.perf_counter()
    orc.probit_elbo_with_grads(t, noise, 0.5, 10.0)
    return time.perf_counter() - t0


def cpu_rows_for(L, Z, S, want=None):
    if want:
        return want
    # pairwise ranking loss materialises ~6 (S,rows,L,L) fp32 tensors (fwd+bwd): keep under ~24 GB and ~10-30 s
    per_row = 6 * S * L * L * 4
    return max(1, min(128, int(24e9 // per_row)))


def run_reference(a):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sh, L, Z, B, S = workload(a)
    torch.set_num_threads(os.cpu_count())
    rows = cpu_rows_for(L, Z, S, a.cpu_rows)
    steps = max(1, a.steps)
    # bound the whole run to a few minutes: probe one step, then cap the step count
    t_probe = cpu_reference_step(L, Z, S, rows)
    budget = 150.0
    steps = max(1, min(steps, int(budget // max(t_probe, 1e-3))))
    warm = max(0, min(a.warmup, 1 if t_probe > 5 else 3))
    for i in range(warm):
        cpu_reference_step(L, Z, S, rows, seed=i)
    ts = [cpu_reference_step(L, Z, S, rows, seed=10 + i) for i in range(steps)]
    t = sum(ts) / len(ts)
    value = S * rows * L / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{sh.name}-shaped S{S} B{B} L{L} Z{Z} fwd+bwd", "S": S, "B_per_gpu": B, "L": L, "Z": Z,
                   "note": "CPU arm runs a bounded row sample of this workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{rows} of {B} rows per step (the reference's O(L^2) ranking loss needs "
                                   f"{6 * S * L * L * 4 / 1e9:.1f} GB per row), oracle port of mpvae.py:145-210 on torch CPU"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from mpvae_b200 import _lib, synth
    from mpvae_b200 import mpvae as M
    from mpvae_b200.probit import contract_nt, contract_tn
    from oracle import probit_elbo_oracle as orc   # only for make_args (a SimpleNamespace) and the CPU baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sh, L, Z, B, S = workload(a)
    Bg = B * world
    flags = {"auto": 0, "fma": _lib.FLAG_CONTRACT_FMA, "tensor": _lib.FLAG_CONTRACT_TENSOR}[a.engine]
    infer = sh.mode == "test"       # BASELINE configs[2] (nuswide) is the test-time path: forward only, no_grad, S = n_test_sample
    args = orc.make_args(L, Z, n_train_sample=S, n_test_sample=S, mode=sh.mode, nll_coeff=0.5, c_coeff=10.0,
                         mpvae_flags=flags, noise_seed=1234, dp_global_batch=Bg, dp_row0=rank * B)

    inp = synth.loss_inputs(L, Z, B, S, seed=100 + rank, label_rate=sh.label_rate, with_noise=False)
    row_keys = ["y", "fe_out", "fe_mu", "fe_logvar", "fx_out", "fx_mu", "fx_logvar"]
    host = {k: torch.from_numpy(inp[k]).pin_memory() for k in row_keys}
    devt = {k: host[k].to(dev) for k in row_keys}
    # the {0,1} label matrix crosses PCIe as bytes (a quarter of the fp32 size); compute_loss casts it on the device
    host["y"] = torch.from_numpy(inp["y"]).to(torch.uint8).pin_memory()
    r_sqrt_sigma = torch.from_numpy(synth.loss_inputs(L, Z, 1, 1, seed=100, with_noise=False)["r_sqrt_sigma"]).to(dev)
    r32 = r_sqrt_sigma.float().requires_grad_(True)      # fp32 working copy of the (replicated) parameter
    flush = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    step_no = [0]
    # the path's one exchange: g_R summed over ranks.  Default: inside the backward over NVLink peer memory
    # (mpvae_b200.peer.PeerRing; measured 34 / 203 us against NCCL's 54 / 270 us for 4 / 64 MB on 8 GPUs,
    # profiles/r01_peer_allreduce.txt); --exchange nccl = an NCCL all-reduce after the backward.  If any rank fails
    # to set the ring up (CUDA IPC unavailable), every rank falls back to NCCL.
    ring = None
    if world > 1 and not infer and a.exchange in ("peer", "nvls"):
        from mpvae_b200.peer import NvlsRing, PeerRing
        try:
            # nvls: the chunk owners reduce inside the NVSwitch (multimem.ld_reduce / multimem.st)
            ring = (NvlsRing if a.exchange == "nvls" else PeerRing)(L, Z, dev)
        except Exception as e:          # noqa: BLE001 -- any failure means "use NCCL", decided collectively below
            print(f"[bench] rank {rank}: peer ring unavailable ({e}); NCCL exchange", file=sys.stderr)
            ring = None
        ok = torch.tensor([1 if ring is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and ring is not None:
            ring = None                 # (its buffers stay allocated; the process is short-lived)
        args.peer_ring = ring

    def one_step(src, from_host):
        args.noise_offset = step_no[0]
        step_no[0] += 1
        t = src
        if infer:
            with torch.no_grad():
                out = M.compute_loss(t["y"], t["fe_out"], t["fe_mu"], t["fe_logvar"], t["fx_out"], t["fx_mu"],
                                     t["fx_logvar"], r32, args)
            if from_host:
                loss_host.copy_(out[0], non_blocking=True)
            return out
        leaves = {k: (t[k] if k == "y" else t[k].requires_grad_(True)) for k in row_keys}
        r32.grad = None
        out = M.compute_loss(leaves["y"], leaves["fe_out"], leaves["fe_mu"], leaves["fe_logvar"], leaves["fx_out"],
                             leaves["fx_mu"], leaves["fx_logvar"], r32, args)
        out[0].backward()
        if world > 1:
            if ring is None:
                dist.all_reduce(r32.grad, op=dist.ReduceOp.AVG)   # the path's one exchange step: NCCL mean over NVLink
            else:
                r32.grad.div_(world)             # with the peer ring the backward already returned the sum
        if from_host:
            loss_host.copy_(out[0].detach(), non_blocking=True)
        for k in row_keys:
            if k != "y":
                t[k].grad = None
                t[k].requires_grad_(False)
        return out

    def timed_e2e(n_steps):
        """K steps through the public API starting from pinned HOST buffers: every step's inputs cross PCIe inside the
        timed region (double-buffered on a side stream by mpvae_b200.train.HostBatchPrefetcher) and the loss comes
        back to the host.  One event pair around all K steps; the 0.8 GB working set is far larger than L2."""
        from mpvae_b200.train import HostBatchPrefetcher
        pf = HostBatchPrefetcher(dev, depth=2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pf.push(host)
        for i in range(n_steps):
            batch = pf.next()
            if i + 1 < n_steps:
                pf.push(host)
            one_step(batch, True)
            pf.release()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def timed(n_steps, from_host):
        evs = []
        for _ in range(n_steps):
            flush.add_(1.0)                         # evict L2 between timed iterations (2 x 126 MB written)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_step(devt, from_host)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [e0.elapsed_time(e1) for e0, e1 in evs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()             # started ahead of the warm-up: its first sample takes a few hundred ms
    for _ in range(max(3, a.warmup)):
        one_step(devt, False)
    timed_e2e(2)
    barrier()
    sampler.mark_begin()            # clocks are reported for the two timed regions below only
    launches0 = _lib.launch_count()
    ms = timed(a.steps, False)
    launches = _lib.launch_count() - launches0
    barrier()
    total_ms = sum(ms)
    total_e2e = timed_e2e(a.steps)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms, total_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, total_e2e = t.tolist()

    # ---- dominant kernel, timed alone on its stream with CUDA events (roofline.achieved) ----
    roof = None
    if rank == 0:
        peaks = measured_peaks()
        M_rows = S * B
        noise = torch.randn(M_rows, Z, device=dev).half().float()     # on the fp16 grid, like the library's Philox noise
        dense = Z >= 128 and L >= 128 and a.engine != "fma"
        from mpvae_b200.probit import contract_workspace
        reps = 5
        if dense:
            # engine 4 prepares the operand planes, engine 5 re-runs the tcgen05 GEMM kernel alone on them: that launch
            # (noise as ONE fp16 piece, R as hi|lo: two MMA passes) is the dominant kernel of the Philox-noise step.
            # MPVAE_TC_CTA=1 has no single-piece variant: engines 2/3, three passes.
            passes = 2 if os.environ.get("MPVAE_TC_CTA", "2") != "1" else 3
            e_prep, e_run = (4, 5) if passes == 2 else (2, 3)
            wsk = contract_workspace(M_rows, L, Z, dev, 2)
            # pitched = rows of the output padded to 16 bytes, exactly as the loss step stores noise.R^T
            contract_nt(noise, r32.detach(), engine=e_prep, ws=wsk, pitched=True)
            run = lambda: contract_nt(noise, r32.detach(), engine=e_run, ws=wsk, pitched=True)
        else:
            eng = {"auto": 0, "fma": 1, "tensor": 2}[a.engine]
            run = lambda: contract_nt(noise, r32.detach(), engine=eng)
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t_k = statistics.median(ts) * 1e-3
        flops = 2.0 * M_rows * L * Z
        if dense:
            kind = os.environ.get("MPVAE_TC_KIND", "f16")
            # split-precision product = `passes` tensor-core passes; fp16 pieces run at the bf16 rate, tf32 at half of it
            div = float(passes) if kind != "tf32" else 2.0 * passes
            peak = peaks["bf16_tflops"] / div
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
            # (profiles/r01_final_ncu_full_summary.txt: 620.58 MB + 148.60 MB per launch) -- eurlex default only
            traffic = 769.18e6 if (L, Z, M_rows, passes, kind) == (3993, 3993, 10240, 2, "f16") else None
            roof = {"bound": "tensor", "achieved": flops / t_k / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": flops / t_k / 1e12 / peak, "traffic": traffic,
                    "kernel": "gemm_split_2sm_kernel<nt> (noise.R^T, mpvae.py:168), tcgen05 " + kind + f" hi/lo split, {passes} MMA passes",
                    "kernel_ms": t_k * 1e3,
                    "peak_basis": f"{peaks['_source']} bf16 burst {peaks['bf16_tflops']} TFLOP/s / {div:g} "
                                  f"(fp32-equivalent flops of a {passes}-pass split-precision product)",
                    "algorithmic_flops_per_launch": flops, "raw_tensor_tflops": passes * flops / t_k / 1e12}
        else:
            bytes_alg = 4.0 * (M_rows * Z + L * Z + M_rows * L)
            roof = {"bound": "hbm", "achieved": bytes_alg / t_k / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": bytes_alg / t_k / 1e9 / peaks["hbm_gbs"], "traffic": None,
                    "kernel": "contract_nt (noise.R^T, mpvae.py:168)", "kernel_ms": t_k * 1e3,
                    "peak_basis": f"{peaks['_source']} HBM copy bandwidth", "algorithmic_bytes_per_launch": bytes_alg}

    # ---- full training step (train.py:103-129: VAE fwd -> loss -> bwd -> all-reduce -> clip -> Adam -> StepLR) ----
    train = None
    if not a.no_train_step and not infer:
        import numpy as np
        from types import SimpleNamespace
        from mpvae_b200.train import DataParallelStep
        margs = SimpleNamespace(feature_dim=sh.feature_dim, label_dim=L, latent_dim=50, z_dim=Z, keep_prob=0.5,
                                scale_coeff=1.0, residue_sigma="", n_train_sample=S, n_test_sample=S, mode="train",
                                nll_coeff=0.5, c_coeff=10.0, mpvae_flags=flags, noise_seed=99)
        np.random.seed(4)
        torch.manual_seed(0)
        vae = M.VAE(margs).to(dev)
        from mpvae_b200.optim import FusedAdam
        opt = FusedAdam(vae.parameters(), lr=torch.tensor(1e-3, device=dev), weight_decay=1e-5)
        sched = torch.optim.lr_scheduler.StepLR(opt, 1000, 0.5)
        stepper = DataParallelStep(vae, opt, sched, margs, clip_norm=100.0)
        rng = np.random.RandomState(5)
        xg = torch.from_numpy(synth.features(Bg, sh.feature_dim, rng)).to(dev)
        yg = torch.from_numpy(synth.labels(Bg, L, sh.label_rate, rng)).to(dev)
        k_train = max(3, min(a.steps, 20))

        def time_steps(fn):
            for _ in range(3):
                fn(yg, xg)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_wall = time.perf_counter()
            e0.record()
            for _ in range(k_train):
                out = fn(yg, xg)
            e1.record()
            torch.cuda.synchronize()
            t_wall = time.perf_counter() - t_wall
            t_dev = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
            return t_dev.item() / k_train, t_wall * 1e3 / k_train, out

        eager_ms, eager_wall, out_t = time_steps(stepper.step)
        train = {"steps_per_s": 1e3 / eager_ms, "ms_per_step": eager_ms, "wall_ms_per_step": eager_wall,
                 "steps": k_train, "global_batch": Bg, "params": int(sum(p.numel() for p in vae.parameters())),
                 "loss": float(out_t.total_loss),
                 "what": "zero_grad, VAE fwd (torch/cuBLAS), probit ELBO fwd+bwd (this library), MLP bwd, "
                         "grad all-reduce, clip_grad_norm_(100) + Adam(wd=1e-5) (this library: mpvae_b200.optim.FusedAdam), "
                         "StepLR; per-step host metrics excluded"}
        try:   # the same step captured once as a CUDA graph and replayed (mpvae_b200.train.GraphedTrainStep)
            if world > 1:
                raise NotImplementedError("graph capture is single-process only")
            from mpvae_b200.train import GraphedTrainStep
            graphed = GraphedTrainStep(stepper)
            g_ms, g_wall, out_g = time_steps(graphed.step)
            train["cuda_graph"] = {"steps_per_s": 1e3 / g_ms, "ms_per_step": g_ms, "wall_ms_per_step": g_wall,
                                   "loss": float(out_g.total_loss)}
        except Exception as exc:   # noqa: BLE001 - the graph leg is informative, never fatal for the bench line
            train["cuda_graph"] = {"error": repr(exc)[:200]}
        del vae, opt, stepper

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    units = float(S) * Bg * L * a.steps
    bi = sum(host[k].numel() * host[k].element_size() for k in row_keys)
    line = {
        "metric": METRIC, "value": units / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{sh.name}-shaped S{S} B{B} L{L} Z{Z} " + ("inference (forward, no_grad)" if infer else "fwd+bwd"),
                   "S": S, "B_per_gpu": B, "B_global": Bg,
                   "L": L, "Z": Z, "D": 50, "noise": "philox (on device, inside the step)", "engine": a.engine,
                   "l2": "L2 flushed between timed iterations (252 MB written); per-step CUDA events summed",
                   "exchange": ("none (1 GPU)" if world == 1 else
                                ("g_R summed inside the NVSwitch (multimem.ld_reduce / multimem.st by the chunk owners), inside "
                                 "the backward") if (ring is not None and a.exchange == "nvls") else
                                "g_R summed over NVLink peer memory inside the backward (chunk owners pull, add in rank "
                                "order, store to every rank)" if ring is not None else "NCCL all-reduce of g_R (fp32) per step")},
        "clocks": clocks,
        "e2e": {"value": units / (total_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": 4,
                "how": "mpvae_b200.compute_loss + backward from pinned host buffers (labels as uint8, the rest fp32); H2D "
                       "double-buffered on a side stream",
                "ms_per_step": total_e2e / a.steps},
        "gpu_launches": int(launches),
        "loss_steps_per_s": a.steps / (total_ms * 1e-3),
        "train_step": train,
        "roofline": roof,
    }
    if not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        rows = cpu_rows_for(L, Z, S, a.cpu_rows)
        cpu_reference_step(L, Z, S, min(rows, 1))
        t_cpu = cpu_reference_step(L, Z, S, rows)
        line["cpu_baseline"] = {"value": S * rows * L / t_cpu, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"1 step on {rows} of {B} rows ({t_cpu:.1f} s): oracle port of mpvae.py:145-210 "
                                          "(O(L^2) pairwise ranking loss) on torch CPU, fwd+bwd"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
