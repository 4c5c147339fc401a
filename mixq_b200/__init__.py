"""mixq_b200 — the MixQ mixed-precision quantized Linear path on B200.

The reference's public names (mixquant/__init__.py:1, mixquant/Cache.py:5) are importable from here; importing the package
does not touch CUDA or the shared library.
"""


def __getattr__(name):
    if name == "AutoForCausalLM":
        from .auto import AutoForCausalLM
        return AutoForCausalLM
    if name == "MixLibCache":
        from .cache import MixLibCache
        return MixLibCache
    raise AttributeError(name)
