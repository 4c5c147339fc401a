"""Print which C-ABI entry points one steady-state decode step of the tiny Llama calls (debug aid)."""
import sys
sys.path.insert(0, ".")
import torch
from mixq_b200 import _lib
from mixq_b200.llama import CONFIGS, LlamaDecoder

lib = _lib.load()
cfg = CONFIGS["tiny"]
m = LlamaDecoder(cfg, batch=24, bit=8, seed=3, outlier_frac=0.02)
tok = torch.randint(0, cfg.vocab, (24, 1)).cuda()
print("discovered:", m.discover(tok))
calls = []
for name in _lib.SIGNATURES:
    fn = getattr(lib, name)
    def wrap(*a, _fn=fn, _n=name):
        calls.append(_n)
        return _fn(*a)
    setattr(lib, name, wrap)
n0 = lib.mixq_launch_count()
m.step(tok)
print("launches", lib.mixq_launch_count() - n0)
print([c for c in calls if c not in ("mixq_launch_count", "mixq_last_error")])
for L in m.layers:
    print({k: (v._n_ind, v.add_outliers, v.forward_without_precondition_len) for k, v in L.items() if hasattr(v, "_n_ind")})
