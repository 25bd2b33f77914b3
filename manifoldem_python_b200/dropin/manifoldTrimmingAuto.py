"""sys.path shim: put this directory BEFORE ManifoldEM's modules/ on sys.path and the reference's
`import manifoldTrimmingAuto` resolves to the B200 implementation (same callables, same config module `p`).
See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)

from manifoldem_python_b200.manifoldTrimmingAuto import *      # noqa: F401,F403,E402
from manifoldem_python_b200.manifoldTrimmingAuto import op     # noqa: F401,E402
