"""One steady-state decode step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from mixq_b200.llama import CONFIGS, LlamaDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-2-7b")
ap.add_argument("--layers", type=int, default=None)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--bit", type=int, default=8)
args = ap.parse_args()
cfg = CONFIGS[args.model]
m = LlamaDecoder(cfg, batch=args.batch, bit=args.bit, layers=args.layers)
tok = torch.randint(0, cfg.vocab, (args.batch, 1)).cuda()
assert m.discover(tok)
m.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.step(tok)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("outliers layer0:", {k: v._n_ind for k, v in m.layers[0].items() if hasattr(v, "_n_ind")})
