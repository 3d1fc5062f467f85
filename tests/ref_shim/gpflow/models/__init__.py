from . import model
