"""One C2 training step (and a few decode steps) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum ...   (launch list)
   ncu --profile-from-start off --set full -k regex:<kernel> ...       (one kernel in depth)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodalanalytical_b200.trainer import FusedTrainer  # noqa: E402
from multimodalanalytical_b200.wrapper import HFWrapper  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c = dict(bench.C2)
model = HFWrapper(data_config=bench.data_config(c), target_tokenizer=bench.Tok(c["V"]), num_steps=100, precision="bf16",
                  seed=bench.SEED, **bench.model_kwargs(c))
if what == "train":
    tr = FusedTrainer(model)
    batch = bench.map_batch(bench.synth_batch(c, B, 1), lambda x: x.cuda())
    for i in range(3):
        tr.train_step(batch, i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.train_step(batch, 3)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    model.eval()
    batch = bench.map_batch(bench.synth_batch(c, B, 1), lambda x: x.cuda())
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    model.generator.generate(*model._relayout(batch, False), n_beams=10, max_length=L, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.generator.generate(*model._relayout(batch, False), n_beams=10, max_length=L, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
