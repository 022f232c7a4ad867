class Path:
    def __init__(self, *a, **k):
        raise NotImplementedError
