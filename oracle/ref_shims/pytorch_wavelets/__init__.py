"""TEST INFRASTRUCTURE ONLY — `pytorch_wavelets` for the REFERENCE arm of tests/test_callers_gpu.py.

The real package is an un-vendored, unpinned dependency of the reference that is not installable here (SURVEY.md
§8c); the reference arm must not touch this repo's kernels, so it gets the plain-PyTorch restatement of the
package's Haar/symmetric transform (oracle/dwt_oracle.py) under the package's name.  Put
`oracle/ref_shims` on PYTHONPATH together with `baseline/_ref/site`; never together with the product package.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)
from oracle.dwt_oracle import DWTForward  # noqa: E402,F401
