BACKENDS_MAPPING = {}


def is_xformers_available():
    return False
