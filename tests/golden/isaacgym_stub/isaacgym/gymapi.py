"""Empty shell: the golden generator never creates a simulator."""
UP_AXIS_Z = 1
IMAGE_COLOR = 0
DOF_MODE_NONE = 0
DOF_MODE_POS = 1


def acquire_gym():
    return None
