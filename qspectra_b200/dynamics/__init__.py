from .base import DynamicalModel, SystemOperator
from .liouville_space import LiouvilleSpaceModel, LiouvilleSpaceOperator
from .redfield import RedfieldModel
from .unitary import UnitaryModel
from .heom import HEOMModel
from .zofe import ZOFEModel
