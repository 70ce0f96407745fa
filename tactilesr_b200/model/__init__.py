from .tactileSR_model import TactileSR, TactileSRCNN, MSRB, ResBlock, Leaky_Res_Block  # noqa: F401
from .tPSFNet import tPSFNet  # noqa: F401
