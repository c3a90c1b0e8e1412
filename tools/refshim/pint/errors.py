class DimensionalityError(Exception):
    pass
