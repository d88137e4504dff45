"""Beam-10 (and greedy) decode throughput over the batch size (SURVEY §8d, C5 sweep) on one GPU."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodalanalytical_b200.wrapper import HFWrapper  # noqa: E402

c = dict(bench.C2)
model = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100, precision="bf16",
                  seed=bench.SEED, **bench.model_kwargs(c))
model.eval()
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 8, 32, 128, 256, 512, 1024]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for B in sizes:
    batch = bench.map_batch(bench.synth_batch(c, B, bench.SEED + 7), lambda x: x.cuda())
    model.generate(batch, n_beams=K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = model.generate(batch, n_beams=K)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(json.dumps({"batch": B, "beams": K, "steps": int(out.shape[1]) - 1, "ms_per_batch": t * 1e3,
                      "ms_per_step": t * 1e3 / max(1, int(out.shape[1]) - 1), "molecules_per_s": B / t}), flush=True)
    model.generator._graphs.clear()
    model.generator._states.clear()
    model.engine._bufs = {k: v for k, v in model.engine._bufs.items() if not str(k[0]).startswith("g.")}
    torch.cuda.empty_cache()
