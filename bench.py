#!/usr/bin/env python
"""Headline benchmark: impressions/sec scored (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of scripts/eval.py's FinalAttentionComponent.transform over the rank's shard
of a MIND-large-shaped synthetic evaluation set (BASELINE.json configs[3]):
  per step = FinalAttention per-row transform of the WHOLE table (5 tcgen05 GEMMs, hoisted from the
             reference's per-history-slot MLPs -- it is redone every step so that no reference
             work is skipped) + ONE fused gather/pool/cosine/dense-rank launch over all impressions.
`value`  : impressions/s with the table, weights and CSR indices already resident in HBM.
`e2e`    : the same step through the public host API (ScoringEngine) with HOST buffers: pinned
           fp32 table + CSR indices copied H2D, scores + ranks copied D2H, every step.
`roofline`: the fused score/rank kernel against the measured HBM copy bandwidth.
`stage_a`: latent-attention pooling (BASELINE.json configs[2] shape, a bounded chunk) as news/s and
           fraction of the measured bf16 tensor peak.
Impressions are independent, so ranks shard them with no data-path collective (weak scaling: every
rank scores its own --impressions; the table is replicated).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from news_recommendation_project_v2_b200 import synthetic as syn  # noqa: E402

METRIC = "impressions_per_sec_scored"
UNIT = "impressions/s"

# MIND-large-shaped workload (SURVEY.md 8d cfg 4)
N_ROWS, DIM, HIDDEN, H_MAX = 161_013, 1024, 4096, 50


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--impressions", type=int, default=2_400_000, help="impressions per rank per step")
    ap.add_argument("--precision", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--no-stage-a", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--only-stage-a", action="store_true", help="profiling runs only: latent-attention pooling leg")
    ap.add_argument("--ref-sample", type=int, default=1024,
                    help="impressions per step of the CPU reference arm (cpu_baseline of the native arm: 4x, once)")
    ap.add_argument("--workload", choices=["cfg4", "cfg5"], default="cfg4",
                    help="cfg4 = headline (replicated table); cfg5 = long-history stress, row-sharded table")
    ap.add_argument("--table-rows", type=int, default=10_000_000, help="cfg5: total table rows")
    ap.add_argument("--gather", choices=["p2p", "dma", "nccl", "none"], default="dma",
                    help="cfg5: all-gather implementation (none = transform only, for diagnosis: results invalid)")
    ap.add_argument("--chunk-rows", type=int, default=32768, help="cfg5: table rows transformed + pushed per chunk")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_device_impressions(n_imp: int, n_rows: int, h_max: int, seed: int, device):
    """cfg-4 shaped CSR impressions generated on the device: H ~ clip(Geom(1/32),1,h_max),
    C ~ clip(round(LogNormal(ln 30, 0.7)), 2, 300), uniform row ids."""
    g = torch.Generator(device=device).manual_seed(seed)
    hist_len = torch.empty(n_imp, device=device).geometric_(1.0 / 32.0, generator=g).clamp_(1, h_max).to(torch.int64)
    cand_len = torch.empty(n_imp, device=device).log_normal_(float(np.log(30.0)), 0.7, generator=g).round_() \
        .clamp_(2, 300).to(torch.int64)
    h_off = torch.zeros(n_imp + 1, dtype=torch.int64, device=device)
    c_off = torch.zeros(n_imp + 1, dtype=torch.int64, device=device)
    torch.cumsum(hist_len, 0, out=h_off[1:])
    torch.cumsum(cand_len, 0, out=c_off[1:])
    n_h, n_c = int(h_off[-1]), int(c_off[-1])
    hist_idx = torch.randint(0, n_rows, (n_h,), generator=g, device=device, dtype=torch.int32)
    cand_idx = torch.randint(0, n_rows, (n_c,), generator=g, device=device, dtype=torch.int32)
    return hist_idx, h_off, cand_idx, c_off, hist_len.to(torch.int32), cand_len.to(torch.int32), n_h, n_c


def make_device_labels(c_off: torch.Tensor, seed: int, device) -> torch.Tensor:
    """int8 click labels per candidate: one uniformly placed positive per impression, the others
    Bernoulli(0.04), at least one negative (evaluation.py:49 needs both classes)."""
    g = torch.Generator(device=device).manual_seed(seed)
    n_imp = c_off.numel() - 1
    cnt = c_off[1:] - c_off[:-1]
    n_c = int(c_off[-1])
    lab = (torch.rand(n_c, generator=g, device=device) < 0.04).to(torch.int8)
    pos = c_off[:-1] + (torch.rand(n_imp, generator=g, device=device, dtype=torch.float64) * cnt.double()).long() \
        .minimum(cnt - 1)
    lab[pos] = 1
    seg = torch.repeat_interleave(torch.arange(n_imp, device=device), cnt)
    tot = torch.zeros(n_imp, dtype=torch.int64, device=device).index_add_(0, seg, lab.long())
    full = torch.nonzero(tot == cnt).flatten()  # all positive: clear the slot after the forced positive
    if full.numel():
        nxt = c_off[:-1][full] + (pos[full] - c_off[:-1][full] + 1) % cnt[full]
        lab[nxt] = 0
    return lab


def run_native(args):
    import torch.distributed as dist

    from news_recommendation_project_v2_b200 import _lib, ops
    from news_recommendation_project_v2_b200.engine import ScoringEngine
    from news_recommendation_project_v2_b200.modeling_utils import FinalAttention

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device(dev)
    if world > 1:
        # keep stdout to the single JSON line (NCCL_DEBUG=VERSION prints its banner on stdout)
        os.environ["NCCL_DEBUG"] = os.environ.get("NRB200_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = measured_peaks()
    if args.only_stage_a:
        print(json.dumps({"stage_a": bench_stage_a(dev, peaks, args)}), flush=True)
        return
    dtype = torch.bfloat16 if args.precision == "bf16" else torch.float32
    es = 2 if dtype == torch.bfloat16 else 4

    # ---- synthetic inputs (random-init weights of the reference architecture) -------------------
    torch.manual_seed(1234)
    model = FinalAttention(DIM, HIDDEN, precision=args.precision).eval()
    model.load_state_dict(syn.make_final_attention_state_dict(DIM, HIDDEN, seed=1234))
    model.to(dev)  # the reference's factory does .to(DEVICE) (modeling_utils.py:274-279)
    table_host = syn.make_table(N_ROWS, DIM, seed=1234).pin_memory()  # fp32, L2-normalised (save_emb.py)
    n_imp = args.impressions
    hist_idx, h_off, cand_idx, c_off, hist_len, cand_len, n_h, n_c = make_device_impressions(
        n_imp, N_ROWS, H_MAX, 1234 + rank, dev)

    eng = ScoringEngine(table_host, model, precision=args.precision, device=dev)
    hist_src = eng.cand
    scores = torch.empty(n_c, dtype=torch.float32, device=dev)
    ranks = torch.empty(n_c, dtype=torch.int32, device=dev)
    flag = ops.new_err_flag(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(kernel_events=None):
        eng.prepare_user_encoder(hist_src)  # dense per-row transform (tcgen05)
        if kernel_events is not None:
            kernel_events[0].record()
        eng.score_device(hist_idx, h_off, cand_idx, c_off, n_c, want_ranks=True, err_flag=flag,
                         out_scores=scores, out_ranks=ranks)
        if kernel_events is not None:
            kernel_events[1].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.nrb_kernel_launches()
    kev = [(ev(), ev()) for _ in range(args.steps)]
    t0, t1 = ev(), ev()
    barrier()
    t0.record()
    for i in range(args.steps):
        step(kev[i])
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    launches = lib.nrb_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ops.raise_on_index_error(flag, "bench")
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_imp * world / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (fused score/rank): algorithmic bytes / launch ---------
    r = 2  # FinalAttention reads x and exp(logit) per history slot
    alg_bytes = (r * n_h + n_c) * DIM * es + 4 * (n_h + n_c) + 8 * n_c
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "score_rank_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "score_rank_kernel", "achieved": round(achieved, 1),
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                "traffic": traffic, "peak_source": peaks["source"], "kernel_ms": round(kernel_ms, 3),
                "algorithmic_bytes_per_launch": int(alg_bytes),
                "share_of_step": round(kernel_ms / ms_step, 4),
                "note": "peak = measured COPY bandwidth (half reads, half writes); this kernel is 99.8 % reads and "
                        "~3 % of its algorithmic bytes are L2 hits, so frac can exceed 1"}

    # ---- end to end through the public API with host buffers -----------------------------------
    pin = lambda x: x.cpu().pin_memory()
    hi_h, hl_h, ci_h, cl_h = pin(hist_idx), pin(hist_len), pin(cand_idx), pin(cand_len)
    ho_h, co_h = pin(h_off), pin(c_off)
    scores_h = torch.empty(n_c, dtype=torch.float32).pin_memory()
    ranks_h = torch.empty(n_c, dtype=torch.int32).pin_memory()
    h2d = table_host.numel() * 4 + (n_h + n_c) * 4 + 2 * (n_imp + 1) * 8
    d2h = n_c * 8

    def e2e_step():
        # public host API: streamed table upload + row transform, then chunk-pipelined H2D | score | D2H
        e = ScoringEngine(table_host, model, precision=args.precision, device=dev, cache_table=False)
        e.score_host(hi_h, ho_h, ci_h, co_h, scores_out=scores_h, ranks_out=ranks_h, n_chunks=8)

    for _ in range(0 if args.no_e2e else 2):
        e2e_step()
    barrier()
    e_steps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    t0.record()
    for _ in range(e_steps):
        e2e_step()
    t1.record()
    barrier()
    te = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_imp * world / (float(te.item()) / e_steps * 1e-3) if e_steps else 0.0
    if e_steps:
        assert torch.equal(scores_h, scores.cpu()) and torch.equal(ranks_h, ranks.cpu()), \
            "e2e and resident paths differ"

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "MIND-large-shaped eval (BASELINE configs[3]): FinalAttention user encoder, "
                               "gather+pool+cosine+dense-rank",
                   "impressions_per_gpu": n_imp, "table_rows": N_ROWS, "dim": DIM, "hidden": HIDDEN,
                   "history_max": H_MAX, "sum_history": n_h, "sum_candidates": n_c,
                   "sharding": "impressions sharded, table replicated, no data-path collective",
                   "l2": "inputs_exceed_l2 (tables %.2f GB + indices %.2f GB per step)" %
                         (3 * N_ROWS * DIM * es / 1e9, 4 * (n_h + n_c) / 1e9)},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
    }

    if rank == 0 and not args.no_stage_a:
        out["stage_a"] = bench_stage_a(dev, peaks, args)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(4 * args.ref_sample, steps=1)  # ~10 s of host work
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_stage_a(dev, peaks, args):
    """configs[2] shape (seq 64, d 768, 512 latents, bf16): a 32,768-item chunk of the 1M-item job."""
    from news_recommendation_project_v2_b200 import config as nrb_config
    from news_recommendation_project_v2_b200.latent_attention import LatentAttentionModel

    d, L, S, items = 768, 512, 64, 32768
    m = LatentAttentionModel(dim=d, num_latents=L, precision="bf16").eval()
    m.load_state_dict(syn.make_latent_state_dict(d, L, seed=1234))
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(items, S, d, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    lens = torch.randint(8, S + 1, (items,), generator=g, device=dev)
    mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).to(torch.int32)
    valid = int(mask.sum())
    old = nrb_config.LATENT_MAX_TOKENS
    nrb_config.LATENT_MAX_TOKENS = int(os.environ.get("NRB200_BENCH_STAGE_A_TOKENS", "262144"))
    try:
        for _ in range(2):
            out = m(x, mask)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        t0.record()
        for _ in range(reps):
            out = m(x, mask)
        t1.record()
        torch.cuda.synchronize()
    finally:
        nrb_config.LATENT_MAX_TOKENS = old
    ms = t0.elapsed_time(t1) / reps
    inner = 8 * 512
    f_ref = 4 * d * inner + 4 * inner * L + 24 * d * d  # reference formulation, K/V projection excluded
    f_exec = 4 * 8 * d * L + 24 * d * d  # executed (Wq.K^T and V.Wout folded)
    tf_ref = valid * f_ref / (ms * 1e-3) / 1e12
    tf_exec = valid * f_exec / (ms * 1e-3) / 1e12
    return {"workload": "latent-attention pooling, %d items x %d tokens, d=%d, L=%d, bf16 (BASELINE configs[2] chunk)"
                        % (items, S, d, L),
            "news_per_s": round(items / (ms * 1e-3), 1), "valid_tokens": valid, "ms": round(ms, 3),
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peaks["bf16_tflops_sustained"],
                         "achieved_reference_flops": round(tf_ref, 1), "achieved_executed_flops": round(tf_exec, 1),
                         "frac": round(tf_ref / peaks["bf16_tflops_sustained"], 4),
                         "frac_executed": round(tf_exec / peaks["bf16_tflops_sustained"], 4),
                         "peak_source": peaks["source"]},
            "finite": bool(torch.isfinite(out).all())}


# ------------------------------------------------------------------------------------------------
def _cpu_reference_step(sd, table, imp):
    """The reference's CPU path restated by the oracle: padded gather -> FinalAttention per history slot
    -> per-impression cosine loop -> per-impression dense rank (oracle.final_second_attention_score)."""
    from oracle import oracle  # cpu_baseline / --impl reference are the two places bench may run it

    return oracle.final_second_attention_score(sd, table, imp.hist_idx, imp.hist_len, imp.cand_idx, imp.cand_len,
                                               dtype=torch.float32, batch=64)


def cpu_baseline(sample: int, steps: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.make_final_attention_state_dict(DIM, HIDDEN, seed=1234)
    table = syn.make_table(N_ROWS, DIM, seed=1234)
    imp = syn.make_impressions(sample, N_ROWS, h_max=H_MAX, cand="large", seed=1234)
    _cpu_reference_step(sd, table, syn.make_impressions(64, N_ROWS, h_max=H_MAX, cand="large", seed=1))  # thread pool warm-up
    t = time.perf_counter()
    for _ in range(steps):
        _cpu_reference_step(sd, table, imp)
    dt = (time.perf_counter() - t) / steps
    return {"value": round(sample / dt, 2), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d cfg-4-shaped impressions (d=%d, H<=%d, C~37), fp32 torch CPU ops, oracle port of the "
                      "reference path (per-slot FinalAttention MLP + per-impression cosine + dense rank)"
                      % (sample, DIM, H_MAX)}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port -- the Python reference cannot travel to
    the GPU box) on the host cores, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.make_final_attention_state_dict(DIM, HIDDEN, seed=1234)
    table = syn.make_table(N_ROWS, DIM, seed=1234)
    imp = syn.make_impressions(args.ref_sample, N_ROWS, h_max=H_MAX, cand="large", seed=1234)
    for _ in range(args.warmup):
        _cpu_reference_step(sd, table, imp)
    t = time.perf_counter()
    for _ in range(args.steps):
        _cpu_reference_step(sd, table, imp)
    dt = (time.perf_counter() - t) / args.steps
    value = args.ref_sample / dt
    sample = ("%d cfg-4-shaped impressions per step (d=%d, H<=%d, C~37), fp32, oracle port of the reference "
              "CPU path, %d threads" % (args.ref_sample, DIM, H_MAX, cores))
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MIND-large-shaped eval (BASELINE configs[3]): FinalAttention user encoder, "
                               "gather+pool+cosine+dense-rank", "table_rows": N_ROWS, "dim": DIM, "hidden": HIDDEN,
                   "history_max": H_MAX, "impressions_per_step": args.ref_sample},
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_sharded(args):
    """BASELINE configs[4]: history <= 200, d=1024, 1024 latents (latent-attention user encoder), table
    row-sharded over the ranks: step = per-shard row transform + peer-store all-gather + local scoring."""
    import torch.distributed as dist

    from news_recommendation_project_v2_b200 import _lib, ops
    from news_recommendation_project_v2_b200.latent_attention import LatentAttentionModel
    from news_recommendation_project_v2_b200.sharded import ShardedTableEngine
    from news_recommendation_project_v2_b200.sharding import table_shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device(dev)
    os.environ["NCCL_DEBUG"] = os.environ.get("NRB200_NCCL_DEBUG", "WARN")  # keep stdout to the JSON line
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            os.environ.setdefault("RANK", "0")
            os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = measured_peaks()
    d, L, h_max, n_rows = 1024, 1024, 200, args.table_rows
    n_imp = max(1, 1_048_576 // world) if args.impressions == 2_400_000 else args.impressions
    model = LatentAttentionModel(dim=d, num_latents=L, precision="bf16").eval()
    model.load_state_dict(syn.make_latent_state_dict(d, L, seed=1234))
    model.to(dev)
    r0, r1 = table_shard_bounds(n_rows, world)[rank]
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    local_rows = torch.nn.functional.normalize(
        torch.randn(r1 - r0, d, generator=g, device=dev, dtype=torch.float32), dim=-1).to(torch.bfloat16)
    hist_idx, h_off, cand_idx, c_off, _, _, n_h, n_c = make_device_impressions(n_imp, n_rows, h_max, 1234 + rank, dev)
    eng = ShardedTableEngine(local_rows, n_rows, model, precision="bf16", device=dev, gather=args.gather,
                             chunk_rows=args.chunk_rows)
    scores = torch.empty(n_c, dtype=torch.float32, device=dev)
    ranks = torch.empty(n_c, dtype=torch.int32, device=dev)
    flag = ops.new_err_flag(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(evs=None):
        if evs:
            evs[0].record()
        eng.build(local_rows)
        if evs:
            evs[1].record()
        eng.score_device(hist_idx, h_off, cand_idx, c_off, n_c, want_ranks=True, err_flag=flag, out_scores=scores,
                         out_ranks=ranks)
        if evs:
            evs[2].record()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    launches0 = lib.nrb_kernel_launches()
    kev = [(ev(), ev(), ev()) for _ in range(args.steps)]
    t0, t1 = ev(), ev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0.record()
    for i in range(args.steps):
        step(kev[i])
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.nrb_kernel_launches() - launches0
    ops.raise_on_index_error(flag, "bench")
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    build_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in kev]))
    score_ms = float(np.mean([b.elapsed_time(c) for _, b, c in kev]))
    f_row = 4 * 8 * d * L + 24 * d * d  # executed FLOPs per row (folded heads)
    f_row_ref = 4 * d * 4096 + 4 * 4096 * L + 24 * d * d
    alg_bytes = (n_h + n_c) * d * 2 + 4 * (n_h + n_c) + 8 * n_c
    out = {
        "metric": METRIC, "value": round(n_imp * world / (ms_step * 1e-3), 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "long-history stress (BASELINE configs[4]): latent-attention user encoder, table "
                               "row-sharded, per-shard transform + all-gather (%s) + gather/pool/cosine/rank" % args.gather,
                   "table_rows": n_rows, "rows_per_gpu": r1 - r0, "impressions_per_gpu": n_imp, "dim": d, "latents": L,
                   "history_max": h_max, "sum_history": n_h, "sum_candidates": n_c, "chunk_rows": args.chunk_rows},
        "gpu_launches": int(launches),
        "build": {"ms": round(build_ms, 3),
                  "tflops_executed": round((r1 - r0) * f_row / (build_ms * 1e-3) / 1e12, 1),
                  "tflops_reference_formulation": round((r1 - r0) * f_row_ref / (build_ms * 1e-3) / 1e12, 1),
                  "allgather_bytes_in_per_gpu": int(2 * (n_rows - (r1 - r0)) * d * 2)},
        "roofline": {"bound": "hbm", "kernel": "score_rank_kernel", "achieved": round(alg_bytes / (score_ms * 1e-3) / 1e9, 1),
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(alg_bytes / (score_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                     "traffic": None, "kernel_ms": round(score_ms, 3)},
        "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_sharded(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
