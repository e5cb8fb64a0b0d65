"""Phase / per-kernel-family timing of one workload step with CUDA events (tuning aid, not the bench)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from u2tokenizer_b200 import _lib, ops  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--per-op", action="store_true")
    a = ap.parse_args()
    cfg, geom, spec = bench.make_geometry(a.workload)
    model = bench.build_model(cfg, geom)
    eng = model.engine()
    from u2tokenizer_b200.synthetic import synthetic_inputs
    images, ids, qids = synthetic_inputs(geom, batch=spec["batch"], frames=spec["frames"], n_question=spec["n_question"], lt=spec["lt"])
    images, ids, qids = images.cuda(), ids.cuda(), qids.cuda()
    B, C = images.shape[:2]
    fr = images.reshape(B * C, 1, *images.shape[2:])
    t_enc, feats = timeit(lambda: eng.encode_images(fr))
    v_tokens = feats.view(B, C, feats.shape[-2], feats.shape[-1])
    t_tokens = ops.embed_splice(qids, eng.embed, None)
    t_tok, vis = timeit(lambda: eng.u2tokenizer(v_tokens, t_tokens))
    emb = ops.embed_splice(ids, eng.embed, vis)
    t_pre, hidden = timeit(lambda: eng.prefill(emb))
    print(f"encode_images {t_enc:.2f} ms | u2tokenizer {t_tok:.2f} ms | prefill(L={emb.shape[1]}) {t_pre:.2f} ms", flush=True)
    if spec["mode"] == "generate":
        nt = spec["new_tokens"]
        t_gen, _ = timeit(lambda: eng.generate_greedy(emb, max_new_tokens=nt), n=2)
        print(f"generate_greedy({nt}) {t_gen:.1f} ms -> decode {(t_gen - t_pre) / (nt - 1):.3f} ms/step (graph)", flush=True)
        cache = eng.new_cache(B, emb.shape[1] + 8)
        eng.prefill(emb, cache)
        t_step, _ = timeit(lambda: eng.decode_step(cache), n=4)
        print(f"decode_step eager {t_step:.3f} ms", flush=True)
    if a.per_op:
        # per-entry-point device time via a profiler-free trick: wrap _lib.check to bracket each call with events
        import collections
        acc = collections.defaultdict(lambda: [0.0, 0])
        orig = {}
        lib = _lib.load()
        for name in _lib.SIGNATURES:
            if name in ("u2_version", "u2_last_error", "u2_device_sm_count"):
                continue
            fn = getattr(lib, name)
            orig[name] = fn

            def make(fn, name):
                def wrapped(*args):
                    e0, e1 = ev(), ev()
                    e0.record()
                    rc = fn(*args)
                    e1.record()
                    acc[name][1] += 1
                    acc[name].append((e0, e1))
                    return rc
                return wrapped
            setattr(lib, name, make(fn, name))
        eng.encode_images(fr)
        torch.cuda.synchronize()
        report("encode_images", acc)
        eng.u2tokenizer(v_tokens, t_tokens)
        torch.cuda.synchronize()
        report("u2tokenizer", acc)
        eng.prefill(emb)
        torch.cuda.synchronize()
        report("prefill", acc)
        if spec["mode"] == "generate":
            cache = eng.new_cache(B, emb.shape[1] + 8)
            eng.prefill(emb, cache)
            torch.cuda.synchronize()
            acc.clear()
            eng.decode_step(cache)
            torch.cuda.synchronize()
            report("decode_step", acc)


def report(title, acc):
    rows = []
    for name, v in acc.items():
        ms = sum(e0.elapsed_time(e1) for e0, e1 in v[2:])
        rows.append((ms, name, v[1]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"--- {title}: {tot:.2f} ms (sum of per-call event spans)")
    for ms, name, n in rows:
        print(f"   {name:32s} {ms:9.3f} ms  {n:5d} calls  {100 * ms / max(tot, 1e-9):5.1f}%")
    acc.clear()


if __name__ == "__main__":
    main()
