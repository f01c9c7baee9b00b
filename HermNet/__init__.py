"""Drop-in alias: ``import HermNet`` / ``from HermNet.hermnet import HVNet`` resolve to the B200-native package, so
callers written against ``thu-wangz17/HermNet`` (``/root/reference/HermNet/__init__.py:1-4``) run unchanged."""
import sys

import hermnet_b200
from hermnet_b200 import *  # noqa: F401,F403
from hermnet_b200 import data, hermnet, rmnet, utils

for _name, _mod in (("data", data), ("hermnet", hermnet), ("rmnet", rmnet), ("utils", utils)):
    sys.modules[__name__ + "." + _name] = _mod
