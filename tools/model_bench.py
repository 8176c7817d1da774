#!/usr/bin/env python
"""Model-level numbers (BASELINE configs 2 / 5) outside bench.py: python tools/model_bench.py [train|infer|cpu] [level] ..."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harness import workloads as W  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "train"
level = sys.argv[2] if len(sys.argv) > 2 else "dropin"
if what == "cpu":
    print(json.dumps(W.cpu_reference_step()))
    sys.exit(0)
graphs = "graphs" in sys.argv
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if what == "train":
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    print(json.dumps(W.train_bench(dev, per_gpu_batch=b, level=level, steps=5, warmup=3, graphs=graphs)))
else:
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    print(json.dumps(W.infer_bench(dev, per_gpu_batch=b, level=level, steps=3, warmup=2, graphs=graphs)))
