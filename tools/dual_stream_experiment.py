"""Experiment: does running two half-batches of the step on two streams (one CUDA graph with two parallel branches) beat one
full batch?  The GroupNorm / split kernels are HBM-bound and the GEMMs tensor-bound, so the halves could overlap.

    python tools/dual_stream_experiment.py [B]"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sgam_neurips22_b200 import synthetic  # noqa: E402
from sgam_neurips22_b200.model import VQModel  # noqa: E402


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    return bench.time_steps(fn, steps, 1) / steps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    ds = "clevr-infinite"
    model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to(dev).eval()
    full = bench.StepHarness(model, ds, 256, B, 100, dev, use_graph=True)
    ms_full = timed(full.run)
    print(f"one stream,  {B} trajectories per launch: {ms_full:.3f} ms/step  {B / ms_full * 1e3:.1f} frames/s", flush=True)
    for parts in (2, 4):
        if B % parts:
            continue
        hs = [bench.StepHarness(model, ds, 256, B // parts, 200 + i, dev, use_graph=False) for i in range(parts)]
        streams = [torch.cuda.Stream() for _ in hs]

        def both():
            cur = torch.cuda.current_stream()
            for h, s in zip(hs, streams):
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    h.resident()
            for s in streams:
                cur.wait_stream(s)
        both()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            both()
        ms = timed(g.replay)
        print(f"{parts} streams, {B // parts} trajectories each (one graph, parallel branches): {ms:.3f} ms/step  {B / ms * 1e3:.1f} frames/s", flush=True)
        # the same sub-batches one after the other on ONE stream (what the split costs without any overlap)
        def serial():
            for h in hs:
                h.resident()
        g2 = torch.cuda.CUDAGraph()
        serial()
        torch.cuda.synchronize()
        with torch.cuda.graph(g2):
            serial()
        ms2 = timed(g2.replay)
        print(f"1 stream, {parts} x {B // parts} trajectories back to back: {ms2:.3f} ms/step  {B / ms2 * 1e3:.1f} frames/s", flush=True)
        del hs, g, g2


if __name__ == "__main__":
    main()
