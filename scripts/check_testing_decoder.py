"""The TEST-ONLY library's decoder (instrumented kernels) through forward-only and forward + gradient launches on many tiles per CTA,
against the product library bit for bit -- run before trusting a timeline trace (a kernel variant that hangs shows up here under a
short timeout instead of inside a long GPU job)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hortimapping_b200 import _testing                      # noqa: E402
from hortimapping_b200.decoder import Decoder               # noqa: E402
from tests.helpers import pepper_weights                    # noqa: E402
from scripts.probe_decoder import make_rows, calibrated     # noqa: E402
W, b, codes = pepper_weights()
t = torch.from_numpy(make_rows(codes, int(os.environ.get("N_ROWS", "131072")))).cuda()
ref, dec = calibrated(Decoder(W, b), codes), calibrated(_testing.testing_decoder(W, b), codes)
for jac in (False, True):
    a = dec._eval_rows(t, with_jac=jac)
    torch.cuda.synchronize()
    r = ref._eval_rows(t, with_jac=jac)
    same = torch.equal(a[0], r[0]) and (not jac or torch.equal(a[1], r[1]))
    print("testing library,", "forward + gradient" if jac else "forward", "ok, identical to the product:", same, flush=True)
    assert same
