"""ORACLE -- test infrastructure only (see oracle/README.md).  Never imported by spartan_b200/."""
from . import extent, distarray, expr            # noqa: F401
from . import views                               # noqa: F401  (installs Expr.__getitem__ / transpose / reshape)
from . import builders                            # noqa: F401  (eye / identity / diag* / std into expr)
from . import fio                                 # noqa: F401
from .distarray import initialize                 # noqa: F401
