class Stationary(object):
    pass
