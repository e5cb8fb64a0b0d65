"""A/B of decode-step launch variants at the bench configuration (cfg 3: mu2-Qwen3-8B, 4 sequences, 288-token prompt,
256 new tokens): every variant runs generate_greedy twice on the same prompt embeddings (first call captures the decode
graph, second call is timed with CUDA events) and its token ids are compared with the first variant's.
usage: python tools/decode_ab.py [variant ...]   (variant = name:key=value,key=value; keys: split, attn_pdl, l2_next, l2_la,
pre, fine). A 128-thread form of the split-KV attention (CTAs that fit on an SM beside a resident decode-linear CTA, so the
next chained launch could prefetch its weights during the attention) was measured with this script and dropped: 4.32 ms
per step against 4.04 ms (profiles/r2_decode_ab_attn_cta_warps.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

DEFAULT = ["base:", "s2:split=2", "s8:split=8", "pdl:attn_pdl=1", "next:l2_next=20", "fine:fine=1"]


def main():
    variants = sys.argv[1:] or DEFAULT
    wl = os.environ.get("U2_PROBE_WORKLOAD", "cfg3")
    cfg, geom, spec = bench.make_geometry(wl)
    model = bench.build_model(cfg, geom)
    eng = model.engine()
    B, n_new = spec["batch"], int(os.environ.get("U2_PROBE_NEW", "256"))
    L = geom.num_3d_query_token + spec["n_question"]
    gen = torch.Generator(device="cuda").manual_seed(7)
    emb = (torch.randn(B, L, geom.hidden_size, device="cuda", generator=gen) * 0.02).bfloat16()
    base = dict(split="auto", attn_pdl=int(eng.attn_pdl), l2_next=eng.l2_next_units, l2_la=eng.l2_lookahead_units,
                pre=eng.pre_stages, fine=int(eng.fine_deps))
    ref_ids, out = None, []
    for v in variants:
        name, _, kv = v.partition(":")
        k = dict(base)
        for item in filter(None, kv.split(",")):
            a, _, b = item.partition("=")
            k[a] = b if a == "split" else int(b)
        eng.attn_pdl, eng.fine_deps = bool(k["attn_pdl"]), bool(k["fine"])
        eng.l2_next_units, eng.l2_lookahead_units, eng.pre_stages = int(k["l2_next"]), int(k["l2_la"]), int(k["pre"])
        os.environ["U2_ATTN_SPLIT"] = str(k["split"])
        eng._gen_state = None  # new cache + new captured graph for this variant
        ids = eng.generate_greedy(emb, n_new)
        torch.cuda.synchronize()
        best = None
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ids = eng.generate_greedy(emb, n_new)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        if ref_ids is None:
            ref_ids = ids.clone()
        same = int((ids == ref_ids).sum()) if ids.shape == ref_ids.shape else -1
        rec = dict(variant=name, knobs=k, ms_per_generate=round(best, 2), ms_per_step=round(best / n_new, 4),
                   tokens_per_s=round(B * n_new / (best / 1e3), 1), ids_equal_to_first=f"{same}/{ref_ids.numel()}")
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", os.environ.get("U2_AB_OUT", "decode_ab.json")), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
