"""Times KV-cached beam decode (ms per step) for a list of (batch, beams); env switches select kernel variants.
    python scripts/decode_bench.py 256x10 1x10 1024x10 [--gated]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodalanalytical_b200.wrapper import HFWrapper

gated = "--gated" in sys.argv
c = dict(bench.C2)
over = bench.PAPER if gated else {}
m = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=10, precision="bf16", seed=1,
              **bench.model_kwargs(c, **over))
m.eval()
out = {}
for spec in [a for a in sys.argv[1:] if "x" in a]:
    B, K = (int(x) for x in spec.split("x"))
    batch = bench.map_batch(bench.synth_batch(c, B, 5), lambda x: x.cuda())
    t, steps = bench.time_generate(m, batch, K, 3)
    out[spec] = round(t * 1e3 / steps, 4)
    m.engine.release_buffers()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("MMA_")}, "gated": gated, "ms_per_step": out}))
