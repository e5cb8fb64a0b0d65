#!/usr/bin/env python
"""Benchmark of the mu2-LLM hot path (driver contract: one JSON line on stdout from rank 0).

A "step" is one pass of the hot path over one batch of synthetic CT volumes: vision front ->
mu2-tokenizer -> splice -> decoder prefill -> greedy decode of `new_tokens` report tokens, called
through the reference-facing HuggingFace-style API (`model.generate(images, input_ids,
question_ids=..., max_new_tokens=..., do_sample=False)`).

  value : volumes/s with the inputs already resident in HBM when the timed region starts
  e2e   : the same call with HOST (pinned) inputs, H2D copies and the D2H read of the ids timed
  roofline : the dominant kernel (decode-step weight-streaming GEMV, HBM-bound) timed live with
             CUDA events in isolation, algorithmic bytes / time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the fp32 oracle port (oracle/u2_oracle.py, a restatement of the
             reference's PyTorch forward) timed on this box's host cores on a bounded sample

Workloads (BASELINE.json configs): cfg3 (default) = mu2-Qwen3-8B greedy generate 256 tokens, batch 4,
256^3 volumes (8 frames); cfg2 = mu2-Qwen3-1.7B forward, one 256x256x128 volume.
Multi-GPU: pure data parallel replicas (independent volumes, no data-path collective), weak scaling.
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_geometry(workload: str):
    from u2tokenizer_b200.configuration import QWEN3_1P7B, QWEN3_8B, U2Qwen3Config
    from u2tokenizer_b200.geometry import Geometry
    if workload == "cfg3":
        cfg = U2Qwen3Config(**QWEN3_8B)
        spec = dict(model="mu2-Qwen3-8B", batch=4, frames=8, new_tokens=256, n_question=32, lt=512, mode="generate")
    elif workload == "cfg2":
        cfg = U2Qwen3Config(**QWEN3_1P7B)
        spec = dict(model="mu2-Qwen3-1.7B", batch=1, frames=4, new_tokens=0, n_question=32, lt=512, mode="forward")
    elif workload == "tiny":  # plumbing check only
        cfg = U2Qwen3Config(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                            num_key_value_heads=2, head_dim=64, vocab_size=1024, image_size=[16, 64, 64],
                            vit_hidden_size=128, vit_mlp_dim=256, vit_num_layers=2, vit_num_heads=4, u2t_num_layers=2,
                            u2t_top_k=16, num_3d_query_token=16, tie_word_embeddings=False)
        spec = dict(model="tiny", batch=2, frames=2, new_tokens=8, n_question=8, lt=16, mode="generate")
    else:
        raise SystemExit(f"unknown workload {workload}")
    return cfg, Geometry.from_hf(cfg), spec


def build_model(cfg, geom, seed=0):
    """Random-init weights of the real architecture, generated on the device (no network for checkpoints)."""
    from u2tokenizer_b200.modeling import U2Qwen3ForCausalLM
    from u2tokenizer_b200.synthetic import synthetic_state_dict
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            model = U2Qwen3ForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    sd = synthetic_state_dict(geom, seed=seed, device="cuda", dtype=torch.bfloat16)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    del sd
    model.eval()
    torch.cuda.empty_cache()
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
TRAFFIC_NCU = {"dlinear_chain": 398770688}  # profiles/r1_ncu_summary.md (r1b_dlinear_chain_full.ncu-rep, mean of 2 launches)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------
# roofline probe: the dominant kernel timed live (CUDA events, kernel launched alone in a loop over
# all decoder layers' weights so the working set (>= 4 GB) is far larger than the 126 MB L2)
# ------------------------------------------------------------------------------------------------
def roofline_probe(model, spec, geom):
    from u2tokenizer_b200 import ops
    eng = model.engine()
    hbm, tf, src = measured_peaks()
    if spec["mode"] == "generate":
        # dominant kernel of the generate workload: the decode-step linear chain launch
        # (dlinear_tcgen05_kernel: o_proj -> gate|up -> down -> next qkv in ONE launch, 386 MB of weights for 8B)
        B = spec["batch"]
        hq, hkv, dh, I, E = (geom.num_attention_heads, geom.num_key_value_heads, geom.head_dim, geom.intermediate_size,
                             geom.hidden_size)
        bufs = eng._decode_buffers(B)
        eng.reset_decode_state(B)
        x, qkv, ctx, act, xg_a, xg_b = (bufs[k] for k in ("x", "qkv", "ctx", "act", "xg", "xg2"))
        ctx.normal_()
        x.normal_()
        c0 = dict(ws=bufs["ws"][0], counters=bufs["counters"][0], sched=eng.dl_sched)
        c1 = dict(ws=bufs["ws"][1], counters=bufs["counters"][1], sched=eng.dl_sched)
        nl = len(eng.layers)

        def chain(li):
            w, wn = eng.layers[li], eng.layers[(li + 1) % nl]
            fl = bufs["flags"][li] if eng.fine_deps else [None] * 4
            dep = lambda i, shift: dict(dep_flags=fl[i], dep_shift=shift) if fl[i] is not None else {}
            return [(ctx, w["wo"], x, dict(residual=x, gamma_next=w["ln2"], xg=xg_a, ssq_out=bufs["ssq_a"], ssq_zero=bufs["ssq_b"], out_flags=fl[0], **c0)),
                    (xg_a, w["wgu"], act, dict(ssq_in=bufs["ssq_a"], eps=geom.rms_norm_eps, silu_pair=True, out_flags=fl[1], **dep(0, 1), **c1)),
                    (act, w["wdown"], x, dict(residual=x, gamma_next=wn["ln1"], xg=xg_b, ssq_out=bufs["ssq_b"], ssq_zero=bufs["ssq_a"], out_flags=fl[2], **dep(1, 0), **c0)),
                    (xg_b, wn["wqkv"], qkv, dict(ssq_in=bufs["ssq_b"], eps=geom.rms_norm_eps, **dep(2, 1), **c1))]

        def sweep():
            bufs["step"] += 1
            for li in range(nl):
                ops.dlinear_multi(chain(li), gridbar=bufs["gridbar"][li * 4:(li + 1) * 4], step_dev=bufs["step"], pdl=eng.pdl)
        st = torch.cuda.current_stream()
        sweep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 3
        e0.record(st)
        for _ in range(reps):
            sweep()
        e1.record(st)
        torch.cuda.synchronize()
        eng.reset_decode_state(B)
        sec = e0.elapsed_time(e1) / 1e3 / (reps * nl)
        nq = (hq + 2 * hkv) * dh
        w_bytes = 2 * (E * hq * dh + 2 * I * E + E * I + nq * E)
        act_bytes = 2 * B * (hq * dh + 3 * E + 2 * E + 2 * I + 2 * E + nq)  # activations in/out of the four linears
        alg_bytes = w_bytes + act_bytes
        ach = alg_bytes / sec / 1e9
        return {"bound": "hbm", "kernel": "dlinear_tcgen05_kernel<128> (decode chain: o_proj+gate|up+down+qkv in one launch)",
                "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4),
                "traffic": TRAFFIC_NCU.get("dlinear_chain"), "peak_source": src, "bytes_per_launch": alg_bytes,
                "us_per_launch": round(sec * 1e6, 2),
                "note": "timed live with CUDA events over all layers' weights (13.9 GB working set >> 126 MB L2)"}
    # forward workloads: the ViT MLP GEMM (largest share of tensor work)
    Fr = spec["batch"] * spec["frames"]
    M = Fr * 2056
    a = torch.randn(M, geom.vit_hidden, device="cuda").bfloat16()
    w = eng.vit[0]["w1"]
    out = torch.empty(M, geom.vit_mlp, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(a, w, eng.vit[0]["b1"], act=ops.ACT_GELU, out=out)
    # the kernel is short (< 100 us): replay 10 launches from a CUDA graph so that the host's launch path is not what
    # gets timed (operands stay L2-resident between launches, as they are inside the model's own launch sequence)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            ops.linear(a, w, eng.vit[0]["b1"], act=ops.ACT_GELU, out=out)
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / 30
    fl = 2.0 * M * geom.vit_mlp * geom.vit_hidden
    ach = fl / sec / 1e12
    return {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (ViT MLP fc1 + GELU)", "achieved": round(ach, 1), "peak": tf,
            "unit": "TFLOP/s", "frac": round(ach / tf, 4), "traffic": None, "peak_source": src,
            "flops_per_launch": fl, "us_per_launch": round(sec * 1e6, 2)}


def extra_rooflines(model, spec, geom):
    """Secondary kernels named by BASELINE.json's north star, timed live (CUDA events, inputs larger than L2):
    the 3-D patch-embed brick gather (HBM) and the decoder prefill gate|up GEMM (tensor pipe)."""
    from u2tokenizer_b200 import ops
    eng = model.engine()
    hbm, tf, src = measured_peaks()
    out = []
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # --- patch-embed gather: fp32 volume -> bf16 patch rows, 32 frames = 268 MB in, 134 MB out per launch
    Fr = 32
    D0, D1, D2 = geom.image_size
    vols = [torch.rand(Fr, D0, D1, D2, device="cuda") for _ in range(2)]
    rows = [torch.empty(Fr * geom.n_patches, geom.patch_dim, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    for i in range(2):
        ops.patchify(vols[i], geom.patch_size, out=rows[i])
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    for r in range(8):
        ops.patchify(vols[r % 2], geom.patch_size, out=rows[r % 2])
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / 8
    by = Fr * D0 * D1 * D2 * 6
    out.append({"kernel": "patchify_tma_kernel (3-D patch-embed brick gather, fp32 volume -> bf16 patch rows)", "bound": "hbm",
                "achieved": round(by / sec / 1e9, 1), "peak": hbm, "unit": "GB/s", "frac": round(by / sec / 1e9 / hbm, 4),
                "bytes_per_launch": by, "us_per_launch": round(sec * 1e6, 2), "peak_source": src})
    del vols, rows
    # --- decoder prefill GEMM (gate|up): M = batch * prompt rows
    M = max(spec["batch"], 1) * (geom.num_3d_query_token + spec["n_question"])
    E, I = geom.hidden_size, geom.intermediate_size
    a = (torch.randn(M, E, device="cuda") * 0.05).bfloat16()
    c = torch.empty(M, 2 * I, device="cuda", dtype=torch.bfloat16)
    nl = len(eng.layers)
    for li in range(min(nl, 4)):
        ops.linear(a, eng.layers[li]["wgu"], out=c)
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    for li in range(nl):
        ops.linear(a, eng.layers[li]["wgu"], out=c)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / nl
    fl = 2.0 * M * 2 * I * E
    out.append({"kernel": f"gemm_bf16_tcgen05_kernel<256> (decoder prefill gate|up, M={M} N={2 * I} K={E})", "bound": "tensor",
                "achieved": round(fl / sec / 1e12, 1), "peak": tf, "unit": "TFLOP/s", "frac": round(fl / sec / 1e12 / tf, 4),
                "flops_per_launch": fl, "us_per_launch": round(sec * 1e6, 2), "peak_source": src})
    return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on host cores, bounded sample, extrapolated by layer counts
# ------------------------------------------------------------------------------------------------
def cpu_baseline(geom, spec, budget_note=True):
    """Times the fp32 oracle (a restatement of the reference's PyTorch forward) on the host cores.
    Bounded sample: ONE volume with one layer of each stack (ViT block, SVR layer, TTA layer, decoder
    layer at prefill and for a few cached decode tokens); the per-layer times are scaled by the real
    layer counts to estimate one full step of the workload."""
    import copy
    from oracle import u2_oracle as O
    from u2tokenizer_b200.synthetic import synthetic_inputs, synthetic_state_dict
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1

    def best_threads(fn):
        """The box may report far more logical CPUs than it can run well: pick the fastest thread count."""
        best, best_t = avail, float("inf")
        for n in sorted({min(avail, c) for c in (4, 8, 16, 32, 64, avail)}):
            torch.set_num_threads(n)
            fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = n, dt
        return best

    a_big, b_big = torch.randn(2048, 768), torch.randn(768, 3072)
    a_vec, b_vec = torch.randn(1, geom.hidden_size), torch.randn(geom.hidden_size, geom.intermediate_size)
    n_big = best_threads(lambda: a_big @ b_big)
    n_vec = best_threads(lambda: a_vec @ b_vec)
    cores = max(n_big, n_vec)
    torch.set_num_threads(n_big)
    g1 = copy.deepcopy(geom)
    g1.vit_layers, g1.u2t_num_layers, g1.num_hidden_layers = 1, 1, 1
    g1.vocab_size = min(geom.vocab_size, 8192)  # lm_head timed separately below at its real size per token
    sd = {k: v.float() for k, v in synthetic_state_dict(g1, seed=0, device="cpu", dtype=torch.bfloat16).items()}
    images, ids, qids = synthetic_inputs(g1, batch=1, frames=spec["frames"], n_question=spec["n_question"], lt=spec["lt"])
    t = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        t[name] = time.perf_counter() - t0
        return r

    with torch.no_grad():
        fr = images.view(spec["frames"], 1, *g1.image_size)
        x = timed("patch_embed", lambda: O.patch_embed(sd, "model.vision_tower.vision_tower.", fr, g1.patch_size))
        x = torch.cat((sd["model.vision_tower.vision_tower.cls_token"].expand(x.shape[0], -1, -1), x), 1)
        x = timed("vit_block", lambda: O.vit_block(sd, "model.vision_tower.vision_tower.blocks.0.", x, g1.vit_heads))
        feats = timed("projector", lambda: O.spatial_pooling_projector(sd, "model.mm_projector.", x[:, 1:], g1))
        v = feats.view(1, spec["frames"], -1, g1.hidden_size)
        txt = torch.nn.functional.embedding(qids, sd["model.embed_tokens.weight"])
        v1 = timed("svr_layer", lambda: O.svr_layer(sd, "model.u2tokenizer.svt_module.attention_network.layers.0.", v,
                                                     g1.u2t_num_heads, g1.attn_type))
        g_sel = copy.deepcopy(g1)
        g_sel.u2t_num_layers = 0
        vis = timed("select_pool", lambda: O.svr(sd, "model.u2tokenizer.svt_module.", v1, g_sel))
        q = sd["model.u2tokenizer.query_tokens"]
        g_t0 = copy.deepcopy(g1)
        timed("tta_layer_plus_linagg", lambda: O.tta(sd, "model.u2tokenizer.tta_module.", q, vis, txt, g_t0))
        g_t0.u2t_num_layers = 0
        timed("linagg", lambda: O.tta(sd, "model.u2tokenizer.tta_module.", q, vis, txt, g_t0))
        L = ids.shape[1]
        emb = torch.randn(1, L, g1.hidden_size) * 0.02
        (_, past) = timed("dec_layer_prefill", lambda: O.decoder_forward(sd, emb, g1, return_hidden=True))
        n_tok = 4
        torch.set_num_threads(n_vec)
        def dec():
            p = past
            for _ in range(n_tok):
                _, p = O.decoder_forward(sd, torch.randn(1, 1, g1.hidden_size) * 0.02, g1, p, return_hidden=True)
        timed("dec_layer_decode4", dec)
        head = torch.randn(min(geom.vocab_size, 32768), g1.hidden_size)
        hx = torch.randn(1, g1.hidden_size)
        timed("lm_head_32k_rows", lambda: hx @ head.t())
    log("[cpu_baseline] parts (s):", {k: round(v, 4) for k, v in t.items()}, "threads gemm/gemv", n_big, n_vec, "of", avail)
    nl_v, nl_u, nl_d = geom.vit_layers, geom.u2t_num_layers, geom.num_hidden_layers
    tta_layer = max(t["tta_layer_plus_linagg"] - t["linagg"], 0.0)
    vision = t["patch_embed"] + nl_v * t["vit_block"] + t["projector"] + nl_u * t["svr_layer"] + t["select_pool"] \
        + nl_u * tta_layer + t["linagg"]
    prefill = nl_d * t["dec_layer_prefill"]
    head_tok = t["lm_head_32k_rows"] * geom.vocab_size / head.shape[0]
    per_tok = nl_d * t["dec_layer_decode4"] / n_tok + head_tok
    per_volume = vision + prefill + spec["new_tokens"] * per_tok + (head_tok if spec["new_tokens"] else head_tok * L)
    vols = 1.0 / per_volume
    sample = (f"oracle port (fp32 torch, {n_big} threads for GEMM phases / {n_vec} for decode, best of a sweep over "
              f"{avail} logical CPUs): 1 volume x {spec['frames']} frames; timed 1 ViT block, 1 SVR layer, "
              f"1 TTA layer, 1 decoder layer (prefill L={L} + {n_tok} cached tokens), lm_head slice; scaled by layer counts "
              f"({nl_v}/{nl_u}/{nl_d}) and {spec['new_tokens']} new tokens; measured {sum(t.values()):.1f} s of CPU work")
    return {"value": vols, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample,
            "per_volume_s": per_volume, "tokens_per_s": (1.0 / per_tok) if spec["new_tokens"] else None,
            "parts_s": {k: round(v, 4) for k, v in t.items()}}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("U2_BENCH_WORKLOAD", "cfg3"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg, geom, spec = make_geometry(args.workload)
    metric = "ct_volumes_per_sec"
    base = {"metric": metric, "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {spec['model']} {spec['mode']}, batch {spec['batch']}/GPU, "
                                   f"{spec['frames']} frames of {'x'.join(map(str, geom.image_size))} per volume, "
                                   f"{spec['new_tokens']} new tokens, prompt {geom.num_3d_query_token + spec['n_question']} "
                                   f"tokens, question pad {spec['lt']}",
                       "batch_per_gpu": spec["batch"], "frames": spec["frames"], "new_tokens": spec["new_tokens"],
                       "parallelism": f"dp{args.gpus} (independent replicas, no data-path collective)",
                       "l2": "weights (>= 3.4 GB) and volumes (67 MB each) exceed the 126 MB L2; no explicit flush"}}

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        vals = []
        cb = None
        for _ in range(max(1, min(args.steps, 2))):
            cb = cpu_baseline(geom, spec)
            vals.append(cb["value"])
        v = statistics.median(vals)
        out = dict(base)
        out.update({"impl": "reference", "value": v, "ms_per_step": 1e3 * spec["batch"] / v, "dtype": "f32",
                    "n_gpus": args.gpus, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "tokens_per_sec": cb["tokens_per_s"], "wall_s": round(time.time() - t0, 1)})
        out["cpu_baseline"]["value"] = v
        print(json.dumps(out), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from u2tokenizer_b200 import _lib, parallel
    from u2tokenizer_b200.synthetic import synthetic_inputs
    log(f"[rank {rank}] building {spec['model']} ...")
    model = build_model(cfg, geom)
    images, ids, qids = synthetic_inputs(geom, batch=spec["batch"], frames=spec["frames"], n_question=spec["n_question"],
                                         lt=spec["lt"], seed=1234 + rank)
    h_images, h_ids, h_q = images.pin_memory(), ids.pin_memory(), qids.pin_memory()
    d_images, d_ids, d_q = h_images.cuda(), h_ids.cuda(), h_q.cuda()
    B = spec["batch"]

    def run(im, i, q):
        if spec["mode"] == "generate":
            return model.generate(im, i, question_ids=q, max_new_tokens=spec["new_tokens"], do_sample=False)
        return model(images=im, input_ids=i, question_ids=q).logits[:, -1].float().argmax(-1)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launches()
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1), device="cuda")  # device time, slowest rank
        return ms, _lib.launches() - n0, r

    for _ in range(max(args.warmup, 3)):
        run(d_images, d_ids, d_q)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof = os.environ.get("U2_PROFILE_TIMED", "0") != "0"  # ncu --profile-from-start off: capture only the timed region
    if prof:
        torch.cuda.profiler.start()
    ms_dev, launches, res = timed(lambda: run(d_images, d_ids, d_q), args.steps)
    if prof:
        torch.cuda.profiler.stop()

    def e2e_step():
        out = run(h_images.cuda(non_blocking=True), h_ids.cuda(non_blocking=True), h_q.cuda(non_blocking=True))
        return out.cpu()
    e2e_step()
    ms_e2e, _, res_h = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    n_new = int(res.shape[1]) if spec["mode"] == "generate" else 0
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    vols = world * B * args.steps
    value = vols / (ms_dev / 1e3)
    e2e_v = vols / (ms_e2e / 1e3)
    out = dict(base)
    out.update({"value": round(value, 4), "ms_per_step": round(ms_dev / args.steps, 3), "dtype": "bf16",
                "tokens_per_sec": round(world * B * n_new * args.steps / (ms_dev / 1e3), 2) if n_new else None,
                "e2e": {"value": round(e2e_v, 4), "unit": "volumes/s",
                        "h2d_bytes_per_step": int(h_images.numel() * 4 + h_ids.numel() * 8 + h_q.numel() * 8),
                        "d2h_bytes_per_step": int(res_h.numel() * res_h.element_size()),
                        "tokens_per_sec": round(world * B * n_new * args.steps / (ms_e2e / 1e3), 2) if n_new else None},
                "gpu_launches": int(launches), "clocks": clocks})
    try:
        out["roofline"] = roofline_probe(model, spec, geom)
    except Exception as e:  # the probe must never cost the bench line
        out["roofline"] = {"error": repr(e)}
    try:
        out["roofline_other_kernels"] = extra_rooflines(model, spec, geom)
    except Exception as e:
        out["roofline_other_kernels"] = [{"error": repr(e)}]
    if world == 1 and not args.no_cpu_baseline:
        try:
            del model
            torch.cuda.empty_cache()
            cb = cpu_baseline(geom, spec)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            out["cpu_baseline"]["tokens_per_s"] = cb["tokens_per_s"]
        except Exception as e:
            out["cpu_baseline"] = {"error": repr(e)}
    if dist is not None:
        dist.destroy_process_group()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
