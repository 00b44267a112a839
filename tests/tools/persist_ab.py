"""A/B measurement of zfb_set_option("fir_persist") on the bench workload (cfg2,
uint8 IQ, 512 frames per step, device-resident), in one process:

    python -m tests.tools.persist_ab [--steps 30] [--frames 512] [--skip-check]

First the bit-exactness check of tests/engine_suite.fast_persistent_fir (uses the
oracle for its golden rows -- this is a measurement tool under tests/, not the
product), then device-timed steps (CUDA events on the engine's stream, inputs
larger than L2) per variant, interleaved A/B/A/B so that clock drift hits all
alike, and the FIR chain's own time per launch from the engine's profile.
One JSON line per variant on stdout.
"""
from __future__ import annotations

import argparse
import json

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--variants", default="0,2,2s,1",
                    help="CTAs per SM of the persistent kernel (0 = off); suffix s = static round robin")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--strips-async", default="1", help="comma list; 0: edge strips on the main stream")
    args = ap.parse_args()

    import torch
    from pypanadapter_b200 import synth
    from pypanadapter_b200.engine import ZoomPSD

    eng = ZoomPSD(0)
    if not args.skip_check:
        from tests import engine_suite as es
        es.fast_persistent_fir(eng)
        print(json.dumps({"check": "fast_persistent_fir", "ok": True}), flush=True)

    w = synth.WORKLOADS["cfg2"]
    F = args.frames
    host = synth.make_frames(w, F, distinct=min(F, 8))
    d_in = torch.from_numpy(host.view(np.uint8).reshape(F, -1)).cuda()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    variants = [(int(v.rstrip("s")), v.endswith("s"), int(sa))
                for sa in args.strips_async.split(",") for v in args.variants.split(",")]
    rows = {}
    best = {v: None for v in variants}
    for rnd in range(args.rounds):
        for v in variants:
            eng.set_option("strips_async", v[2])
            eng.set_option("fir_persist", v[0])
            eng.set_option("fir_persist_static", int(v[1]))
            eng.configure(w.fs, w.fft_size, w.fft_ratio, w.frame_len, w.window, dtype=w.dtype, flip=w.flip,
                          f_demod=w.f_demod, crop=w.crop, ema_alpha=w.ema_alpha, mode="fast")
            d_rows = torch.empty((F, eng.row_width), dtype=torch.float32, device="cuda")
            eng.reset_ema()
            for _ in range(args.warmup):
                eng.process_device(d_in.data_ptr(), F, d_rows.data_ptr())
            torch.cuda.synchronize()
            eng.profile()
            eng.set_profiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                eng.process_device(d_in.data_ptr(), F, d_rows.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            eng.set_profiling(False)
            prof = eng.profile()
            ms = e0.elapsed_time(e1) / args.steps
            fir_ms, fir_n = prof.get("decimate_stage0", (0.0, 1))
            eng.reset_ema()
            eng.process_device(d_in.data_ptr(), F, d_rows.data_ptr())
            torch.cuda.synchronize()
            rows[v] = d_rows.cpu().numpy()
            rec = {"fir_persist": v[0], "static": v[1], "strips_async": v[2], "round": rnd, "ms_per_step": ms,
                   "gsamples_per_s": F * w.frame_len / ms / 1e6,
                   "fir_chain_ms_per_launch": fir_ms / max(fir_n, 1),
                   "rows_equal_to_variant0": bool(np.array_equal(rows[v], rows[variants[0]]))}
            if best[v] is None or ms < best[v]["ms_per_step"]:
                best[v] = rec
            print(json.dumps(rec), flush=True)
    print(json.dumps({"best": [best[v] for v in variants]}), flush=True)


if __name__ == "__main__":
    main()
