"""sys.path shim: with this directory BEFORE ManifoldEM's modules/ on sys.path every `import myio` of the reference
(57 call sites) resolves to the B200 package's record reader / writer — required for the 'sidecar' record layout,
harmless for the default 'pickle' layout (same bytes as modules/myio.py).  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)

from manifoldem_python_b200.myio import fin1, fout1, fout2, Record     # noqa: F401,E402
