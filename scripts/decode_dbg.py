import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multimodalanalytical_b200.wrapper import HFWrapper
from multimodalanalytical_b200.trainer import FusedTrainer
c = dict(bench.C2)
model = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100, precision="bf16", seed=bench.SEED, **bench.model_kwargs(c))
def run(B, tag):
    batch = bench.map_batch(bench.synth_batch(c, B, bench.SEED + 7), lambda x: x.cuda())
    model.eval()
    out = model.generate(batch, n_beams=10)
    torch.cuda.synchronize()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = model.generate(batch, n_beams=10); e1.record(); torch.cuda.synchronize()
        print(tag, B, rep, "ms/step %.3f" % (e0.elapsed_time(e1) / (out.shape[1] - 1)), "steps", out.shape[1] - 1, flush=True)
for B in (64, 256, 64):
    run(B, "fresh")
if len(sys.argv) > 1:
    tr = FusedTrainer(model)
    hb = [bench.map_batch(bench.synth_batch(c, 256, i), lambda x: x.cuda()) for i in range(4)]
    for i in range(30):
        tr.train_step(hb[i % 4], i)
    torch.cuda.synchronize()
    for B in (64, 256, 1, 64, 1024, 64):
        run(B, "after-train")
