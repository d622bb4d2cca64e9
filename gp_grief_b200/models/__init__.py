from .basemodel import BaseModel
from .gp_grief_model import GPGriefModel
from .gp_web_model import GPwebModel
