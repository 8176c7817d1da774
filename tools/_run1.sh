timeout 200 python tools/model_bench.py train fused 24 graphs 2>&1 | tail -2 | cut -c1-400
