"""One training step of a bench.train_case (c2_paper / c3 / c4) between cudaProfilerStart/Stop, for ncu launch lists:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/profile_case.py c4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodalanalytical_b200.trainer import FusedTrainer  # noqa: E402
from multimodalanalytical_b200.wrapper import HFWrapper  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
tc = bench.train_case(name)
model = HFWrapper(data_config=tc["dc"], target_tokenizer=bench.Tok(tc["V"]), num_steps=100, precision="bf16",
                  seed=bench.SEED, **tc["mk"])
tr = FusedTrainer(model, use_graph=False)
batch = bench.map_batch(tc["batch"](tc["B"], 1), lambda x: x.cuda())
for i in range(3):
    tr.train_step(batch, i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.train_step(batch, 3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
