"""Import shim that lets the UNMODIFIED reference run on a CPU-only box (or on the CPU of a GPU box).

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden*.py (run in the build container, where
/root/reference is mounted) to produce the golden vectors under tests/golden/, and by
oracle/ref_runner.py, which bench.py's baseline arms (`--impl reference`, `--impl reference-cuda`, the
`cpu_baseline` leg) use to time the unmodified reference staged under baseline/_ref/
(scripts/vendor_reference.py).  Nothing in the product path imports this file.

What it neutralises:
  * wild_completion/utils.py:14-18 imports addict / plyfile / open3d / skimage at module top
    -> stub modules (only `addict.Dict` is subclassed, utils.py:524);
  * 31 hard-coded `.cuda()` calls (loss.py:33,55,...; utils.py:162,...) -> identity on CPU;
  * deepsdf/deep_sdf/workspace.py:217 `torch.load` without map_location -> map to CPU;
  * utils.get_time (utils.py:614-619) calls torch.cuda.synchronize -> no-op.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def install(reference_root: str = REFERENCE_ROOT, force_cpu: bool = False):
    """force_cpu: neutralise the hard-coded `.cuda()` calls even when a GPU is present (the CPU baseline arm on the GPU box)."""
    import torch

    def _stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class _Dict(dict):
        def __getattr__(self, k):
            return self[k]

        __setattr__ = dict.__setitem__

    if "addict" not in sys.modules:
        _stub("addict", Dict=_Dict)
    if "plyfile" not in sys.modules:
        _stub("plyfile")
    if "open3d" not in sys.modules:
        o3d = _stub("open3d")
        o3d.geometry = types.SimpleNamespace()
        o3d.utility = types.SimpleNamespace(random=types.SimpleNamespace(seed=lambda s: None))
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.measure = _stub("skimage.measure")

    if force_cpu or not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
        _ld = torch.load
        if not getattr(_ld, "_hm_shim", False):
            def _load(f, *a, **k):
                k.setdefault("map_location", "cpu")
                return _ld(f, *a, **k)
            _load._hm_shim = True
            torch.load = _load
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
