def hsv_to_rgb(*a, **k):
    raise NotImplementedError
