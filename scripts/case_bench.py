"""ms/step of bench.train_case configs (graph-replayed FusedTrainer steps): python scripts/case_bench.py c3 c4 ..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
out = bench.bench_train_configs([a for a in sys.argv[1:] if not a.startswith("-")], 20, 1, 0, None)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("MMA_")},
                  "ms": {k: round(v["ms_per_step"], 3) for k, v in out.items()}}))
