"""Drop-in namespace for the parts of ``spconv`` the reference imports (``import spconv.pytorch as spconv``,
``from spconv.utils import Point2VoxelCPU3d``): see pytorch.py / utils.py."""
from . import pytorch  # noqa: F401
from . import utils  # noqa: F401
