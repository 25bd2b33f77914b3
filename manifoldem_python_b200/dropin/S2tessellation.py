"""sys.path shim: `import S2tessellation` (modules/Data.py:6) resolves to the B200 implementation.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.append(_root)

from manifoldem_python_b200.S2tessellation import op, classS2, get_S2, sphere_points     # noqa: F401,E402
