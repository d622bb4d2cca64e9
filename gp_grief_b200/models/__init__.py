from .basemodel import BaseModel
from .gp_grief_model import GPGriefModel
