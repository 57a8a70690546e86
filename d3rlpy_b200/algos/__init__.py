from .base import VectorEncoderFactory  # noqa: F401
from .bcq import BCQ  # noqa: F401
from .cql import CQL  # noqa: F401
from .dqn import DQN, NFQ, DiscreteCQL, DoubleDQN, PixelEncoderFactory, QRQFunctionFactory  # noqa: F401
from .td3_plus_bc import TD3PlusBC  # noqa: F401
from .sac import SAC  # noqa: F401
from .td3 import TD3  # noqa: F401
from .ddpg import DDPG  # noqa: F401
from .iql import IQL  # noqa: F401
from .awac import AWAC  # noqa: F401
from .crr import CRR  # noqa: F401
from .plas import PLAS  # noqa: F401
from .bear import BEAR  # noqa: F401
