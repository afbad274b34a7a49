#!/usr/bin/env python
"""bench.py -- the W4A16 dequant-GEMM hot path of ferrum-infer-rs on B200.

  python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W      # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): the Llama-3.1-8B GPTQ-Int4 dense linears
(qkv 4096->6144, o 4096->4096, gate_up 4096->28672, down 14336->4096; group 128, act-order
g_idx, asymmetric zeros) of all 32 layers, at every decode batch M in {1,2,4,8,16,32}.
One STEP = one pass over those 32*4*6 = 768 GEMMs with distinct random-init packed weights
(3.6 GB of weights per M: far larger than the 126 MB L2, so every GEMM streams from HBM).
`value` = algorithmic bytes of the step (SURVEY.md 8d formula) / device time of the step.
Extra series on the same line: prefill M=512 TFLOP/s, Qwen3-30B-A3B MoE-FFN decode tok/s.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA = {"hidden": 4096, "inter": 14336, "layers": 32,
         "linears": [("qkv_proj", 4096, 6144), ("o_proj", 4096, 4096),
                     ("gate_up_proj", 4096, 28672), ("down_proj", 14336, 4096)]}
QWEN3 = {"hidden": 2048, "inter": 768, "experts": 128, "top_k": 8, "layers": 48}
# Gemma3-27B (crates/ferrum-quantization/tests/gptq_parity_test.rs:584-660): desc_act checkpoint, GELU-tanh gate
GEMMA3 = {"hidden": 5376, "inter": 21504, "q": 4096, "kv": 2048,
          "linears": [("qkv_proj", 5376, 8192), ("o_proj", 4096, 5376), ("gate_up_proj", 5376, 43008),
                      ("down_proj", 21504, 5376)]}
DECODE_MS = [1, 2, 4, 8, 16, 32]
GROUP = 128


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(d["bf16_tflops"]),
                "tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "src": "fallback"}


def alg_bytes(k, n, m, act_order=True):
    g = k // GROUP
    return k * n // 2 + g * n * 2 + g * (n // 8) * 4 + (k * 4 if act_order else 0) + m * k * 2 + m * n * 2


# ------------------------------------------------------------------------ synthetic weights
def synth_gptq(rng, k, n, act_order=True):
    """Random-init packed GPTQ tensors: uniform nibbles, f16-representable scales
    in [0.01, 0.1), zero nibbles 0..14, balanced random g_idx."""
    qw = rng.integers(0, 1 << 32, size=(k // 8, n), dtype=np.uint32).view(np.int32)
    sc = rng.uniform(0.01, 0.1, size=(k // GROUP, n)).astype(np.float16).astype(np.float32)
    zn = rng.integers(0, 15, size=(k // GROUP, n // 8, 8), dtype=np.uint32)
    qz = (zn << (np.arange(8, dtype=np.uint32) * 4)).sum(axis=2, dtype=np.uint32).view(np.int32)
    gi = None
    if act_order:
        gi = np.repeat(np.arange(k // GROUP, dtype=np.int32), GROUP)
        rng.shuffle(gi)
    return qw, sc, qz, gi


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path).read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- GPU arm
def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(val, world):
    import torch
    if world == 1:
        return val
    import torch.distributed as dist
    t = torch.tensor([val], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_graph(graph, steps, warmup, world):
    """W warm-up replays, then exactly K replays between CUDA events on the launching stream."""
    import torch
    for _ in range(warmup):
        graph.replay()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)   # ms for K steps


def capture(fn):
    import torch
    fn()                                   # eager once: grows scratch, sets func attributes
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            fn()
    return g


def run_gpu(args):
    import torch
    fb2 = importlib.import_module("ferrum-infer-rs_b200")
    from importlib import import_module
    par = import_module("ferrum-infer-rs_b200.parallel")
    rank, world, local = dist_setup(args.gpus)
    pk = peaks()
    rng = np.random.default_rng(1234 + 17 * rank)
    layers = args.layers or LLAMA["layers"]
    dev = f"cuda:{local}"

    # ---------------------------------------------------------------- build the linears
    # N > 1: tensor parallel (column split for qkv / gate_up, row split for o / down + all-reduce)
    t_build = time.time()
    lins = []          # [layer][4]
    shapes = []        # per-rank (k, n) of each of the 4 linears
    host_layer0 = None
    wrng = np.random.default_rng(99)    # same full weights on every rank, sharded below
    for li in range(layers):
        row = []
        for (name, k, n) in LLAMA["linears"]:
            qw, sc, qz, gi = synth_gptq(wrng, k, n, act_order=not args.no_act_order)
            if li == 0 and rank == 0 and world == 1:
                host_layer0 = host_layer0 or []
                host_layer0.append((name, k, n, qw, sc, qz, gi))
            if world > 1:
                qw, sc, qz, gi, k, n = par.shard_gptq(name, qw, sc, qz, gi, k, n, GROUP, rank, world)
            row.append(fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, GROUP, k, n, device=local))
            if li == 0:
                shapes.append((k, n))
        lins.append(row)
    build_s = time.time() - t_build
    max_m = max(DECODE_MS + [args.prefill_m])
    fb2.pregrow_scratch(local, max_m)

    full_shapes = [(k, n) for (_, k, n) in LLAMA["linears"]]
    xs = {m: [torch.randn(m, k, device=dev).clamp_(-1, 1).half() for (k, n) in shapes] for m in DECODE_MS}
    ys = {m: [torch.empty(m, n, device=dev, dtype=torch.float16) for (k, n) in shapes] for m in DECODE_MS}
    row_parallel = [False, True, False, True]
    comm = None
    if world > 1:
        # the library's own communicator; decode-sized row-parallel outputs are all-reduced by the
        # fused path (GEMM epilogue pushes tiles to the peers over NVLink + one reduce kernel)
        comm = par.Comm.from_torch_distributed(local)
        comm.enable_p2p(max(DECODE_MS) * max(n for (k, n) in shapes))

    parity = None
    if world > 1 and not args.quick:
        parity = parity_check(fb2, par, rank, world, local, comm)

    def step_for(ms, fused=True):
        def fn():
            for m in ms:
                for li in range(layers):
                    for j in range(4):
                        if world > 1 and row_parallel[j]:
                            if fused:
                                comm.forward_allreduce(lins[li][j], xs[m][j], ys[m][j], m)
                            else:
                                lins[li][j].forward(xs[m][j], ys[m][j], m)
                                par.all_reduce_(ys[m][j])
                        else:
                            lins[li][j].forward(xs[m][j], ys[m][j], m)
        return fn

    # ------------------------------------------------------------------ headline: decode
    fb2.reset_launch_count()
    g_dec = capture(step_for(DECODE_MS))
    launches_per_step = fb2.launch_count() // 2          # eager pass + capture pass
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = time_graph(g_dec, args.steps, args.warmup, world)
    clk = clocks.stop() if rank == 0 else {}
    ms_step = ms_total / args.steps
    n_gemms = layers * 4 * len(DECODE_MS)
    step_bytes = sum(alg_bytes(k, n, m, not args.no_act_order) for m in DECODE_MS for (k, n) in full_shapes) * layers
    value = step_bytes / (ms_step * 1e-3) / 1e9

    # per-M series (same graphs, one M each) and roofline of the dominant kernel
    series = {}
    for m in DECODE_MS:
        g = capture(step_for([m]))
        t = time_graph(g, max(3, args.steps // 2), 2, world) / max(3, args.steps // 2)
        b = sum(alg_bytes(k, n, m, not args.no_act_order) for (k, n) in full_shapes) * layers
        series[f"M={m}"] = {"gbs": round(b / (t * 1e-3) / 1e9, 1), "us_per_gemm": round(t * 1e3 / (layers * 4), 2),
                            "frac_hbm": round(b / (t * 1e-3) / 1e9 / pk["hbm_gbs"], 4)}
        del g
    tp_allreduce = None
    if world > 1:
        # same decode step with the row-parallel all-reduces done by NCCL (torch.distributed) instead
        g = capture(step_for(DECODE_MS, fused=False))
        t_nccl = time_graph(g, max(3, args.steps // 2), 2, world) / max(3, args.steps // 2)
        del g
        tp_allreduce = {"fused_push_ms_per_step": round(ms_step, 4), "nccl_ms_per_step": round(t_nccl, 4),
                        "speedup": round(t_nccl / ms_step, 3),
                        "note": "row-parallel o/down: GEMM epilogue pushes tiles to peer memory + reduce kernel, vs GEMM + ncclAllReduce"}
    per_shape = {}
    if world == 1 and not args.quick:
        for j, (name, k, n) in enumerate(LLAMA["linears"]):
            def fn(j=j):
                for li in range(layers):
                    lins[li][j].forward(xs[1][j], ys[1][j], 1)
            g = capture(fn)
            t = time_graph(g, 5, 2, world) / 5
            b = alg_bytes(k, n, 1, not args.no_act_order) * layers
            per_shape[name] = {"gbs_M1": round(b / (t * 1e-3) / 1e9, 1), "us": round(t * 1e3 / layers, 2)}
            del g

    # -------------------------------------------------------------------------- prefill
    prefill = None
    if not args.quick:
        pm = args.prefill_m
        xp = [torch.randn(pm, k, device=dev).clamp_(-1, 1).half() for (k, n) in shapes]
        yp = [torch.empty(pm, n, device=dev, dtype=torch.float16) for (k, n) in shapes]
        pl = min(layers, 8)

        def fn_p():
            for li in range(pl):
                for j in range(4):
                    lins[li][j].forward(xp[j], yp[j], pm)
                    if world > 1 and row_parallel[j]:
                        par.all_reduce_(yp[j])
        g = capture(fn_p)
        t = time_graph(g, 3, 2, world) / 3
        fl = sum(2 * pm * k * n for (k, n) in full_shapes) * pl
        tf = fl / (t * 1e-3) / 1e12
        prefill = {"M": pm, "tflops": round(tf, 1), "frac_tensor": round(tf / pk["tflops_sustained"], 4),
                   "peak_tflops": pk["tflops_sustained"], "kernel": lins[0][0].kernel_name(pm)}
        del g

    # ------------------------------------------------------------------- e2e (host buffers)
    # Same step through the host-buffer plugin call: H2D of every GEMM's input from pinned host
    # memory and D2H of every output inside the timed region (copies of chunk c+1 / c-1 overlap the
    # GEMMs of chunk c on separate streams), one sync per M.
    e2e = None
    if world == 1:
        flat = [lins[li][j] for li in range(layers) for j in range(4)]
        # ONE call per step: every GEMM of the step with its own row count (the copy pipeline fills and drains once)
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # page-locked: DMA'd directly (pageable ones would be staged)
        all_l, all_x, all_y, all_m = [], [], [], []
        for m in DECODE_MS:
            hx = [pin(np.random.default_rng(5).uniform(-1, 1, size=(m, k)).astype(np.float16)) for (k, n) in shapes] * layers
            hy = [pin(np.empty((m, n), dtype=np.float16)) for li in range(layers) for (k, n) in shapes]
            all_l += flat; all_x += hx; all_y += hy; all_m += [m] * len(flat)
        batches = [fb2.HostBatch(all_l, all_x, all_y, all_m)]
        for _ in range(max(3, args.warmup)):
            for b in batches:
                b.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for b in batches:
                b.run()                      # each call ends with a stream sync + copy-out
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e = {"value": round(step_bytes / (e_ms * 1e-3) / 1e9, 1), "unit": "GB/s",
               "h2d_bytes_per_step": sum(b.h2d_bytes for b in batches),
               "d2h_bytes_per_step": sum(b.d2h_bytes for b in batches),
               "ms_per_step": round(e_ms, 3), "api": "fb2_gptq_linear_forward_host_batch_m (one call per step, cached launch graphs)"}
    else:
        # N > 1: the same experiment as N = 1, per rank: every one of the step's GEMMs gets its input from pinned host
        # memory and returns its (reduced) output there -- 768 H2D and 768 D2H copies per step inside the timed region
        hx = {m: [torch.randn(m, k).half().pin_memory() for (k, n) in shapes] for m in DECODE_MS}
        hy = {m: [torch.empty(m, n, dtype=torch.float16).pin_memory() for (k, n) in shapes] for m in DECODE_MS}

        def e2e_step():
            for m in DECODE_MS:
                for li in range(layers):
                    for j in range(4):
                        xs[m][j].copy_(hx[m][j], non_blocking=True)
                        if row_parallel[j]:
                            comm.forward_allreduce(lins[li][j], xs[m][j], ys[m][j], m)
                        else:
                            lins[li][j].forward(xs[m][j], ys[m][j], m)
                        hy[m][j].copy_(ys[m][j], non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(3):
            e2e_step()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier(world)
        e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, world)
        e2e = {"value": round(step_bytes / (e_ms * 1e-3) / 1e9, 1), "unit": "GB/s",
               "h2d_bytes_per_step": sum(int(t.nbytes) for m in DECODE_MS for t in hx[m]) * layers * world,
               "d2h_bytes_per_step": sum(int(t.nbytes) for m in DECODE_MS for t in hy[m]) * layers * world,
               "ms_per_step": round(e_ms, 3),
               "api": "fb2_gptq_linear_forward / fb2_gptq_linear_forward_allreduce per GEMM, host <-> device copies per GEMM (per rank)"}

    # ------------------------------------------------------- model-level decode step
    model_level = None
    if world == 1 and not args.quick:
        model_level = run_model_level(args, fb2, lins, layers, local, series)

    # ------------------------------------------- Gemma3-27B TP prefill (BASELINE.json configs[4])
    tp_prefill = None
    if not args.quick and not args.no_tp_prefill:
        del lins
        torch.cuda.empty_cache()
        tp_prefill = run_tp_prefill(args, fb2, par, rank, world, local, pk, comm)

    # ---------------------------------------------------------------------- MoE series
    moe = None
    if not args.quick and not args.no_moe:
        moe = run_moe(args, fb2, par, rank, world, local, pk, comm)

    # ----------------------------------------------------------------------- CPU baseline
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and host_layer0 is not None:
        cpu = cpu_baseline(host_layer0, threads=1, ms=DECODE_MS if not args.quick else [1])

    if rank == 0:
        avg_launch_us = ms_step * 1e3 / n_gemms
        achieved = (step_bytes / n_gemms) / (avg_launch_us * 1e-6) / 1e9
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = (f"ncu dram read+write of one {tj.get('workload')} launch "
                            f"(algorithmic {tj.get('algorithmic_bytes_per_launch')} B), {tj.get('source')}")
        line = {
            "metric": "w4a16_gemm_decode_throughput", "value": round(value, 1), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f16 (int4 weights, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "Llama-3.1-8B GPTQ-Int4 dense linears (qkv/o/gate_up/down), decode M=1..32, "
                                   f"{layers} layers of distinct random-init packed weights, group 128, "
                                   f"{'act-order g_idx' if not args.no_act_order else 'sequential groups'}, asymmetric zeros",
                       "gemms_per_step": n_gemms, "decode_M": DECODE_MS, "layers": layers,
                       "l2_defeat": "inputs larger than L2: each GEMM streams distinct weights (3.6 GB per M) from HBM",
                       "parallelism": f"tp{world}" if world > 1 else "single"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": pk["src"],
                         "kernel": "dk_gemm_kernel (decode kernel, M <= 32): stream-K, tcgen05",
                         "avg_launch_us": round(avg_launch_us, 3),
                         "alg_bytes_per_launch": int(step_bytes / n_gemms)},
            "cpu_baseline": cpu, "series_decode": series, "per_shape_M1": per_shape, "prefill": prefill,
            "moe_decode": moe, "tp_prefill_gemma3": tp_prefill, "tp_allreduce": tp_allreduce, "parity": parity,
            "model_level": model_level,
            "build_s": round(build_s, 1),
            "host_cores": os.cpu_count(),
        }
        print(json.dumps(line), flush=True)
    if parity is not None and not parity["ok"]:
        sys.stderr.write(f"bench: multi-GPU parity check failed: {parity}\n")
        sys.stderr.flush()
        os._exit(3)
    if world > 1:
        # leave without the NCCL / CUDA-graph teardown (it can block when graphs hold collectives)
        barrier(world)
        sys.stdout.flush()
        os._exit(0)


def parity_check(fb2, par, rank, world, local, comm):
    """N > 1: the sharded paths against the single-GPU result of the same library on the same inputs (every rank
    holds the full weights for this one layer): (tp) Llama down_proj, row-parallel, act-order, GEMM + NCCL all-reduce;
    (tp_fused) the same through the fused push all-reduce of the decode step; (ep) one Qwen3-30B-A3B MoE layer,
    expert-parallel through fb2_moe_ep_ffn_forward.  Returns max-abs / max-ref per path."""
    import torch
    import torch.distributed as dist
    dev = f"cuda:{local}"
    rng = np.random.default_rng(4321)                     # same on every rank
    name, k, n = LLAMA["linears"][3]
    qw, sc, qz, gi = synth_gptq(rng, k, n, act_order=True)
    full = fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, GROUP, k, n, device=local)
    tpl = par.TensorParallelLinear(name, qw, sc, qz, gi, None, 4, GROUP, k, n, rank, world, device=local)
    out = {}
    cols = torch.from_numpy(np.asarray(tpl.x_cols, dtype=np.int64)).to(dev)
    for m, key in ((4, "tp_m4"), (32, "tp_m32")):
        x = torch.from_numpy(rng.uniform(-1, 1, size=(m, k)).astype(np.float16)).to(dev)
        y_full = full.forward(x).float()
        x_loc = x[:, cols].contiguous()
        y_nccl = tpl.forward(x_loc).float()                                    # GEMM + all-reduce (torch.distributed)
        y_fused = torch.empty((m, n), dtype=torch.float16, device=dev)
        comm.forward_allreduce(tpl.linear, x_loc, y_fused, m)                  # fused push all-reduce (decode sizes)
        torch.cuda.synchronize()
        ref = y_full.abs().max().item()
        out[key] = round((y_nccl - y_full).abs().max().item() / ref, 6)
        out[key.replace("tp_", "tp_fused_")] = round((y_fused.float() - y_full).abs().max().item() / ref, 6)
    del full, tpl
    H, I, E, K = QWEN3["hidden"], QWEN3["inter"], QWEN3["experts"], QWEN3["top_k"]
    gu = [synth_gptq(rng, H, 2 * I, False)[:3] for _ in range(E)]
    dn = [synth_gptq(rng, I, H, False)[:3] for _ in range(E)]
    mk = lambda parts, kk, nn: fb2.GptqExpertStack.load_gptq_stacked([q[0] for q in parts], [q[1] for q in parts],
                                                                    [q[2] for q in parts], None, 4, GROUP, kk, nn, device=local)
    B = 32
    ffn = fb2.moe.MoeFfn(mk(gu, H, 2 * I), mk(dn, I, H), K, True, max_batch=B)
    e_lo, e_hi = rank * E // world, (rank + 1) * E // world
    ep = par.ExpertParallelFfn(mk(gu[e_lo:e_hi], H, 2 * I), mk(dn[e_lo:e_hi], I, H), K, True, B // world, E, rank, world,
                               comm=comm).enable_native()
    fb2.pregrow_scratch(local, B * K)
    xg = torch.from_numpy(rng.uniform(-1, 1, size=(B, H)).astype(np.float16)).to(dev)
    lg = torch.from_numpy(rng.uniform(-2, 2, size=(B, E)).astype(np.float16)).to(dev)
    y_ref = ffn.forward(xg, lg).float()
    b_loc = B // world
    sl = slice(rank * b_loc, (rank + 1) * b_loc)
    y_ep = ep.forward(xg[sl].contiguous(), lg[sl].contiguous()).float()
    torch.cuda.synchronize()
    out["ep_b32"] = round((y_ep - y_ref[sl]).abs().max().item() / max(y_ref.abs().max().item(), 1e-6), 6)
    # worst value over the ranks
    t = torch.tensor([out[k_] for k_ in sorted(out)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {k_: float(v) for k_, v in zip(sorted(out), t.tolist())}
    out["tolerance"] = 2e-3
    out["ok"] = all(v <= 2e-3 for k_, v in out.items() if k_ not in ("tolerance",))
    return out


def run_model_level(args, fb2, lins, layers, local, gemm_series):
    """one CUDA graph of the Llama-3.1-8B decode step through `layers` GPTQ layers with everything but
    attention: add+RMSNorm (storing straight into the qkv / gate|up layers' act-order column order) -> qkv -> attention
    STUB (q is passed through: one copy kernel) -> o_proj -> add+RMSNorm -> gate|up with the SwiGLU epilogue -> down.
    tok/s = M / step time; `gemm_share` = the same GEMMs alone (series_decode) / this step."""
    import torch
    dev = f"cuda:{local}"
    H, I = LLAMA["hidden"], LLAMA["inter"]
    out = {}
    w_norm = torch.ones(H, device=dev, dtype=torch.float16)
    perms = [[lins[li][j].act_order_perm() for j in range(4)] for li in range(layers)]
    for m in (1, 4, 8, 16, 32):
        x = torch.randn(m, H, device=dev).clamp_(-1, 1).half()
        res = torch.zeros(m, H, device=dev, dtype=torch.float16)
        h = torch.empty(m, H, device=dev, dtype=torch.float16)
        qkv = torch.empty(m, 6144, device=dev, dtype=torch.float16)
        attn = torch.empty(m, H, device=dev, dtype=torch.float16)
        o = torch.empty(m, H, device=dev, dtype=torch.float16)
        act = torch.empty(m, I, device=dev, dtype=torch.float16)
        dn = torch.empty(m, H, device=dev, dtype=torch.float16)

        def fn():
            cur = x
            for li in range(layers):
                q_l, o_l, gu_l, d_l = lins[li]
                fb2.fused_add_rms_norm(cur, w_norm, 1e-5, residual=res, out=h, out_perm=perms[li][0])
                (q_l.forward_sorted if perms[li][0] else q_l.forward)(h, qkv, m)
                attn.copy_(qkv[:, :H])                                  # attention stub
                o_l.forward(attn, o, m)
                fb2.fused_add_rms_norm(o, w_norm, 1e-5, residual=res, out=h, out_perm=perms[li][2])
                gu_l.forward_swiglu(h, act, m, x_is_sorted=perms[li][2] is not None)
                d_l.forward(act, dn, m)
                cur = dn
        fb2.reset_launch_count()
        g = capture(fn)
        launches = fb2.launch_count() // 2 + layers                    # + the stub copies (torch)
        t = time_graph(g, 5, 3, 1) / 5                                   # ms per step of `layers` layers
        del g
        t32 = t * LLAMA["layers"] / layers
        gemm_us = gemm_series.get(f"M={m}", {}).get("us_per_gemm")
        out[f"M={m}"] = {"tok_s_no_attention": round(m / (t32 * 1e-3), 1), "ms_per_token_step_32_layers": round(t32, 4),
                         "us_per_layer": round(t * 1e3 / layers, 2), "launches_per_layer": round(launches / layers, 2),
                         "gemm_share": round(4 * gemm_us / (t * 1e3 / layers), 3) if gemm_us else None}
    out["note"] = ("Llama-3.1-8B-GPTQ-Int4 decode, all 32 layers' linears + norms + SwiGLU in one CUDA graph, attention stubbed "
                   "(q passed through), no lm_head / sampling: an upper bound of what this path allows, next to the "
                   "reference's end-to-end serving figures on an RTX 4090 (README.md:96: 122.9 / 324.3 / 640.2 / 745.6 tok/s "
                   "at c = 1 / 4 / 16 / 32, where the Marlin GEMMs are 73 % of the iteration, "
                   "docs/bench/v0.2-cuda/status-2026-05-05.md:120-136)")
    return out


def run_tp_prefill(args, fb2, par, rank, world, local, pk, comm=None):
    """Gemma3-27B GPTQ-Int4 dense linears, prefill M tokens, tensor parallel over `world` GPUs:
    qkv (column) -> o (row + all-reduce) -> gate_up (column) -> GELU-tanh*up (local) -> down (row +
    all-reduce).  Act-order on every layer; row-parallel shards are cut in g_idx-sorted order."""
    import torch
    dev = f"cuda:{local}"
    M, layers = args.tp_prefill_m, args.tp_prefill_layers
    wrng = np.random.default_rng(2727)
    mods = []
    for _ in range(layers):
        row = {}
        for (name, k, n) in GEMMA3["linears"]:
            qw, sc, qz, gi = synth_gptq(wrng, k, n, act_order=not args.no_act_order)
            row[name] = par.TensorParallelLinear(name, qw, sc, qz, gi, None, 4, GROUP, k, n, rank, world, device=local,
                                                 qkv_segments=[GEMMA3["q"], GEMMA3["kv"], GEMMA3["kv"]])
        mods.append(row)
    fb2.pregrow_scratch(local, M)
    H, I, Q = GEMMA3["hidden"], GEMMA3["inter"], GEMMA3["q"]
    x = torch.randn(M, H, device=dev).clamp_(-1, 1).half()
    qkv = torch.empty(M, mods[0]["qkv_proj"].n_local, device=dev, dtype=torch.float16)
    attn = torch.empty(M, Q // world, device=dev, dtype=torch.float16)
    h1 = torch.empty(M, H, device=dev, dtype=torch.float16)
    gu = torch.empty(M, mods[0]["gate_up_proj"].n_local, device=dev, dtype=torch.float16)
    act = torch.empty(M, I // world, device=dev, dtype=torch.float16)
    h2 = torch.empty(M, H, device=dev, dtype=torch.float16)

    def fn():
        for m in mods:
            m["qkv_proj"].forward(x, qkv, M)
            attn.copy_(qkv[:, : Q // world])                   # stand-in for attention: the local q heads
            m["o_proj"].forward(attn, h1, M, comm)             # + all-reduce
            m["gate_up_proj"].forward(h1, gu, M)
            fb2.moe.fused_gelu_tanh_mul_split(gu, act, M, I // world)
            m["down_proj"].forward(act, h2, M, comm)           # + all-reduce
    g = capture(fn)
    t = time_graph(g, 3, 2, world) / 3                         # ms per pass over `layers` layers
    flops = 2.0 * M * sum(k * n for (_, k, n) in GEMMA3["linears"]) * layers
    tf = flops / (t * 1e-3) / 1e12
    del g
    return {"M": M, "layers": layers, "tflops_total": round(tf, 1), "tflops_per_gpu": round(tf / world, 1),
            "frac_tensor_per_gpu": round(tf / world / pk["tflops_sustained"], 4), "ms_per_layer": round(t / layers, 3),
            "parallelism": f"tp{world}",
            "all_reduce": "fb2_gptq_linear_forward_allreduce (ncclAllReduce through the library communicator), fp16 [M, hidden], 2 per layer"}


def run_moe(args, fb2, par, rank, world, local, pk, comm=None):
    """Qwen3-30B-A3B MoE FFN (128 experts, top-8, gate_up 2048->1536, down 768->2048):
    MoE-FFN-only decode tok/s = B / (48 * t_layer).  N > 1: expert parallel."""
    import torch
    H, I, E, K = QWEN3["hidden"], QWEN3["inter"], QWEN3["experts"], QWEN3["top_k"]
    layers = args.moe_layers
    if world > 1:
        args.moe_batches = [b for b in (32, 64, 128, 256) if b % world == 0]     # BASELINE.json configs[3]
    dev = f"cuda:{local}"
    wrng = np.random.default_rng(4242)
    gi_gu = gi_dn = None
    if not args.no_act_order:
        gi_gu = np.repeat(np.arange(H // GROUP, dtype=np.int32), GROUP); wrng.shuffle(gi_gu)
        gi_dn = np.repeat(np.arange(I // GROUP, dtype=np.int32), GROUP); wrng.shuffle(gi_dn)
    e_lo, e_hi = (rank * E // world, (rank + 1) * E // world)
    ffns = []
    for _ in range(layers):
        gu = [synth_gptq(wrng, H, 2 * I, False)[:3] for _ in range(E)]
        dn = [synth_gptq(wrng, I, H, False)[:3] for _ in range(E)]
        gu, dn = gu[e_lo:e_hi], dn[e_lo:e_hi]
        gst = fb2.GptqExpertStack.load_gptq_stacked([g[0] for g in gu], [g[1] for g in gu], [g[2] for g in gu],
                                                    gi_gu, 4, GROUP, H, 2 * I, device=local)
        dst = fb2.GptqExpertStack.load_gptq_stacked([d[0] for d in dn], [d[1] for d in dn], [d[2] for d in dn],
                                                    gi_dn, 4, GROUP, I, H, device=local)
        if world == 1:
            ffns.append(fb2.moe.MoeFfn(gst, dst, K, True, max_batch=max(args.moe_batches)))
        else:
            ffns.append(par.ExpertParallelFfn(gst, dst, K, True, max(args.moe_batches) // world, E, rank, world,
                                              comm=comm).enable_native())
    fb2.pregrow_scratch(local, max(args.moe_batches) * K)          # EP: world * (B/world * K) received rows
    out = {}
    for B in args.moe_batches:
        if B % world:
            continue
        b_loc = B // world                                        # EP: tokens are batch-sharded
        torch.manual_seed(1000 + B)                               # same global batch on every rank
        xg = torch.randn(B, H, device=dev).clamp_(-1, 1).half()
        lgg = (torch.rand(B, E, device=dev) * 4 - 2).half()
        x = xg[rank * b_loc:(rank + 1) * b_loc].contiguous()
        lg = lgg[rank * b_loc:(rank + 1) * b_loc].contiguous()
        y = torch.empty(b_loc, H, device=dev, dtype=torch.float16)

        def fn():
            for f in ffns:
                f.forward(x, lg, y)
        g = capture(fn)              # NCCL all-to-alls are captured with the kernels (no host work per layer)
        t = time_graph(g, 5, 3, world) / 5
        del g
        t_layer = t / layers                                     # ms
        ids = torch.topk(lgg.float(), K, dim=1).indices
        uniq = int(torch.unique(ids).numel())
        bytes_layer = uniq * (1634304 + 817152) + B * H * 2 * 2
        out[f"B={B}"] = {"tok_s_ffn_only": round(B / (QWEN3["layers"] * t_layer * 1e-3), 1),
                         "us_per_layer": round(t_layer * 1e3, 2), "experts_hit": uniq,
                         "gbs": round(bytes_layer / (t_layer * 1e-3) / 1e9, 1),
                         "frac_hbm": round(bytes_layer / (t_layer * 1e-3) / 1e9 / pk["hbm_gbs"] / max(world, 1), 4)}
    out["note"] = f"{layers} distinct layers rotated; attention etc. excluded; {'ep' + str(world) if world > 1 else 'single GPU'}"
    return out


# ----------------------------------------------------------------------- CPU reference arm
def cpu_baseline(host_layer0, threads, ms):
    """The reference's CPU path (oracle port of cpu.rs:2283-2315 + cpu.rs:481-492) on layer 0's four
    linears: load-time dequant (untimed, reported) then CpuGptqLinear::forward per M."""
    import oracle
    t_deq = time.time()
    ws = []
    for (name, k, n, qw, sc, qz, gi) in host_layer0:
        ws.append((k, n, oracle.dequant_gptq(qw, sc, qz, gi, 4, GROUP, k, n)))
    t_deq = time.time() - t_deq
    rng = np.random.default_rng(7)
    total_b, t0 = 0, time.perf_counter()
    for m in ms:
        for (k, n, w) in ws:
            x = rng.uniform(-1, 1, size=(m, k)).astype(np.float16).astype(np.float32)
            oracle.linear_forward(x, w, None, m, threads=threads)
            total_b += alg_bytes(k, n, m, host_layer0[0][6] is not None)
    dt = time.perf_counter() - t0
    return {"value": round(total_b / dt / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"layer 0's 4 linears at M={ms} ({len(ms) * 4} of the step's GEMMs), f32 weights "
                      f"dequantised at load ({t_deq:.1f} s, untimed) as the reference does",
            "seconds": round(dt, 2), "host_cores": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    all_threads = oracle.max_threads()
    rng = np.random.default_rng(99)
    layer0 = []
    for (name, k, n) in LLAMA["linears"]:
        qw, sc, qz, gi = synth_gptq(rng, k, n, act_order=not args.no_act_order)
        layer0.append((name, k, n, qw, sc, qz, gi))
    ws = [(k, n, oracle.dequant_gptq(qw, sc, qz, gi, 4, GROUP, k, n)) for (_, k, n, qw, sc, qz, gi) in layer0]
    xrng = np.random.default_rng(7)
    xs = {m: [xrng.uniform(-1, 1, size=(m, k)).astype(np.float16).astype(np.float32) for (k, n, _) in ws] for m in DECODE_MS}
    b = sum(alg_bytes(k, n, m, not args.no_act_order) for m in DECODE_MS for (k, n, _) in ws)

    def timed(threads, steps, warmup):
        def step():
            for m in DECODE_MS:
                for j, (k, n, w) in enumerate(ws):
                    oracle.linear_forward(xs[m][j], w, None, m, threads=threads)
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return (time.perf_counter() - t0) / steps
    # The reference's Linux CPU GEMM is single-threaded (cpu.rs:481-492): ONE core is the baseline.  A step costs
    # ~8 s on one core, so the step count is bounded to keep the run within a few minutes.
    steps1 = max(1, min(args.steps, 4))
    dt = timed(1, steps1, min(args.warmup, 1))
    v = round(b / dt / 1e9, 4)
    dt_all = timed(all_threads, 2, 1)
    sample = (f"each step = layer 0's 4 linears at M={DECODE_MS} (24 of the 768 GEMMs of a full step; the other "
              f"31 layers repeat the same shapes), oracle port of the reference CPU path, f64 accumulate; {steps1} timed steps")
    print(json.dumps({
        "impl": "reference", "metric": "w4a16_gemm_decode_throughput", "value": v, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": steps1, "warmup": min(args.warmup, 1), "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 accumulate)",
        "data": "synthetic",
        "config": {"workload": "Llama-3.1-8B GPTQ-Int4 dense linears (qkv/o/gate_up/down), decode M=1..32", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(),
                         "all_cores": {"value": round(b / dt_all / 1e9, 4), "unit": "GB/s", "cores": all_threads,
                                       "note": "courtesy figure: the reference itself never threads this GEMM"}},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fb2", choices=["fb2", "reference"])
    ap.add_argument("--layers", type=int, default=0, help="Llama layers to instantiate (default 32)")
    ap.add_argument("--prefill-m", type=int, default=512)
    ap.add_argument("--moe-layers", type=int, default=12)
    ap.add_argument("--moe-batches", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32])
    ap.add_argument("--no-act-order", action="store_true")
    ap.add_argument("--no-moe", action="store_true")
    ap.add_argument("--no-tp-prefill", action="store_true")
    ap.add_argument("--tp-prefill-m", type=int, default=4096)
    ap.add_argument("--tp-prefill-layers", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only (used under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "fb2" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
