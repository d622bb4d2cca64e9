from . import kern
