"""Per-phase time of the one-launch decode step (decode_step.cu): %globaltimer stamps of cluster 0 at every phase boundary.
    MMA_DECODE_PERSIST_DBG=1 python scripts/decode_step_phases.py 1x10 [--gated]"""
import os, sys
os.environ["MMA_DECODE_PERSIST_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodalanalytical_b200.wrapper import HFWrapper

gated = "--gated" in sys.argv
c = dict(bench.C2)
m = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=10, precision="bf16", seed=1,
              **bench.model_kwargs(c, **(bench.PAPER if gated else {})))
m.eval()
names = ["embed"]
for i in range(6):
    names += [f"L{i}.ln+qkv", f"L{i}.self-attn", f"L{i}.out+res", f"L{i}.ln+q", f"L{i}.cross-attn", f"L{i}.out+res",
              f"L{i}.ln+ffn1", f"L{i}.ffn2+res"]
names += ["ln+lm-head"]
for spec in [a for a in sys.argv[1:] if "x" in a]:
    B, K = (int(x) for x in spec.split("x"))
    batch = bench.map_batch(bench.synth_batch(c, B, 5), lambda x: x.cuda())
    m.generate(batch, n_beams=K)  # graph-replayed: the stamps are those of the last step (t = 126)
    torch.cuda.synchronize()
    t = m.generator.dbg_times.cpu().tolist()
    n = len(names)
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(n)]
    print(f"== {spec} gated={gated}: step {sum(d):.1f} us (last step, t = 126)")
    agg = {}
    for nm, x in zip(names, d):
        k = nm.split(".")[-1]
        agg.setdefault(k, []).append(x)
    for k, v in agg.items():
        print(f"  {k:12s} n={len(v):2d}  avg {sum(v) / len(v):6.2f} us  total {sum(v):7.1f} us   first {v[0]:6.2f}")
    m.engine.release_buffers()
