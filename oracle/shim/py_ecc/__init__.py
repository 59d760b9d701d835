"""TEST INFRASTRUCTURE (oracle) -- stand-in for the third-party ``py_ecc`` package
(py-ecc==7.0.1, /root/reference/requirements.txt:14), restated because the
package is neither vendored by the reference nor installable here.  Put
``oracle/shim`` on ``sys.path`` to let the reference's own modules import it.
"""
from . import fields  # noqa: F401
from . import bn128  # noqa: F401
