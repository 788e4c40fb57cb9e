"""Spectral-product timings only (bench.py's spectral_products leg): mag1c rmf / 30 iterations, ratio, SRF."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = sys.argv[:1]
import bench
pk, _ = bench.peaks()
out = bench.spectral_products_bench(torch.device("cuda", 0), pk)
for k, v in out.items():
    print(k, json.dumps(v) if isinstance(v, dict) else v, file=sys.stderr)
