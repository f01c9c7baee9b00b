"""hermnet_b200 -- B200-native (sm_100a) implementation of HermNet's message-passing hot path behind the
reference's Python API (``HVNet`` / ``HPNet`` / ``HTNet``, ``neighbor_search``, ``transform``, ``in_subgraph``,
``virial_calc``).  See DESIGN.md."""
from .data import *      # noqa: F401,F403
from .hermnet import *   # noqa: F401,F403
from .rmnet import *     # noqa: F401,F403
from .utils import *     # noqa: F401,F403

__version__ = "0.1.0"
