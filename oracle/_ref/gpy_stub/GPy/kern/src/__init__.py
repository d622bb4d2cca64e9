from . import stationary
