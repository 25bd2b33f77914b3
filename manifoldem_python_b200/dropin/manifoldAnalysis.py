"""sys.path shim: `import manifoldAnalysis` resolves to the B200 implementation.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)

from manifoldem_python_b200.manifoldAnalysis import op, divide, fileCheck, count    # noqa: F401,E402
