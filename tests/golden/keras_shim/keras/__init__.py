"""keras_shim: see ../README.md.  TEST INFRASTRUCTURE -- eager float64 stand-in for Keras 2.1.4."""
__version__ = '2.1.4-shim'
from . import backend      # noqa: F401
from . import layers       # noqa: F401
from . import models       # noqa: F401
from . import optimizers   # noqa: F401
from . import utils        # noqa: F401
from . import callbacks    # noqa: F401
from . import constraints  # noqa: F401
from . import regularizers  # noqa: F401
from . import losses       # noqa: F401
