"""Empty shell: tensors are plain torch tensors already."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
