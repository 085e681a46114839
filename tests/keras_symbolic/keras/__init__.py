"""keras_symbolic: see ../README.md.  TEST INFRASTRUCTURE -- `keras` as a recording front end of deephar_b200."""
import os

__version__ = '2.1.4-symbolic'
# optimizers / callbacks / regularizers / constraints / losses / utils: the trivial stand-ins of the eager shim
__path__.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'golden', 'keras_shim', 'keras'))

from . import backend    # noqa: E402,F401
from . import layers     # noqa: E402,F401
from . import models     # noqa: E402,F401
