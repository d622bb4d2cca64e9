class Kern(object):
    pass
from . import src
